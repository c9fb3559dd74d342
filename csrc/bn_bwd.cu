// Batch-norm backward (training mode) over NHWC bf16 activations.
//
//   reduce : dgamma[c] = sum g * xhat,  dbeta[c] = sum g
//   apply  : dx = gamma * invstd * (g - dbeta / M - xhat * dgamma / M)
// with g = dy masked by the ReLU that followed the batch norm:
//   relu 0  no ReLU (projection shortcut branch)
//   relu 1  mask from the stored output y (unit with a residual input)
//   relu 2  mask recomputed from x with the forward scale/shift (no residual) - y is not read.
//   relu 3  mask read from the bit mask the forward bn_apply stored (y points at it: one byte
//           per 8 channels instead of 16 - a residual unit's backward reads 1/16 of y's bytes)
// Both kernels are latency bound unless enough bytes are in flight, so every thread owns one
// 8-channel group (per-channel coefficients live in registers) and issues the 16-byte loads of
// FOUR rows before touching any of them.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "ops.h"
#include "ptx.cuh"

namespace tfos {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 4;

__device__ __forceinline__ float2 up2(uint32_t w) {  // packed bf16 pair -> fp32 pair
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
// gradient pair masked by the ReLU that followed the batch norm (modes: see the file header)
__device__ __forceinline__ float2 masked_g(int relu, uint32_t gw, float2 xv, uint32_t yw,
                                           uint32_t mbits, int j, float2 fs, float2 fh) {
  float2 g = up2(gw);
  if (relu == 2) {
    const float2 pre = __ffma2_rn(xv, fs, fh);
    g.x = pre.x > 0.f ? g.x : 0.f;
    g.y = pre.y > 0.f ? g.y : 0.f;
  } else if (relu == 1) {
    const float2 yv = up2(yw);
    g.x = yv.x > 0.f ? g.x : 0.f;
    g.y = yv.y > 0.f ? g.y : 0.f;
  } else if (relu == 3) {
    g.x = ((mbits >> (2 * j)) & 1u) ? g.x : 0.f;
    g.y = ((mbits >> (2 * j + 1)) & 1u) ? g.y : 0.f;
  }
  return g;
}

struct Geo {
  int groups, cg, rl, g_in, r_in;
};
__device__ __forceinline__ Geo geo(int C) {
  Geo g;
  g.groups = C >> 3;
  g.cg = g.groups < kThreads ? g.groups : kThreads;  // channel groups per block pass
  g.rl = kThreads / g.cg;                            // row lanes
  g.g_in = threadIdx.x % g.cg;
  g.r_in = threadIdx.x / g.cg;
  return g;
}

__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const __nv_bfloat16* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ fscale,
                     const float* __restrict__ fshift, long long P, int C, int relu, float* dgamma,
                     float* dbeta, const int rev) {
  const Geo G = geo(C);
  __shared__ float red[2][kThreads][8];
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int g0 = blockIdx.y * G.cg; g0 < G.groups; g0 += gridDim.y * G.cg) {
    const int g = g0 + G.g_in;
    float2 a0[4], a1[4];  // packed fp32x2 accumulators: channels (2j, 2j+1)
#pragma unroll
    for (int j = 0; j < 4; ++j) a0[j] = a1[j] = make_float2(0.f, 0.f);
    if (g < G.groups && G.r_in < G.rl) {
      const int c = g * 8;
      float2 nmu[4], fs[4], fh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        nmu[j] = make_float2(-mean[c + 2 * j], -mean[c + 2 * j + 1]);
        fs[j] = relu == 2 ? make_float2(fscale[c + 2 * j], fscale[c + 2 * j + 1])
                          : make_float2(0.f, 0.f);
        fh[j] = relu == 2 ? make_float2(fshift[c + 2 * j], fshift[c + 2 * j + 1])
                          : make_float2(0.f, 0.f);
      }
      const long long stride = static_cast<long long>(gridDim.x) * G.rl;
      for (long long p = static_cast<long long>(blockIdx.x) * G.rl + G.r_in; p < P;
           p += kRows * stride) {
        uint4 gq[kRows], xq[kRows], yq[kRows];
        uint32_t mq[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const long long pk = p + k * stride;
          const bool in = pk < P;
          const long long o = (rev ? P - 1 - pk : pk) * C + c;
          gq[k] = in ? ld_nc_v4(dy + o) : zero;
          xq[k] = in ? ld_nc_v4(x + o) : zero;
          yq[k] = (in && relu == 1) ? ld_nc_v4(y + o) : zero;
          mq[k] = (in && relu == 3) ? __ldg(reinterpret_cast<const uint8_t*>(y) + (o >> 3)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          const uint32_t gw[4] = {gq[k].x, gq[k].y, gq[k].z, gq[k].w};
          const uint32_t xw[4] = {xq[k].x, xq[k].y, xq[k].z, xq[k].w};
          const uint32_t yw[4] = {yq[k].x, yq[k].y, yq[k].z, yq[k].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 xv = up2(xw[j]);
            const float2 gj = masked_g(relu, gw[j], xv, yw[j], mq[k], j, fs[j], fh[j]);
            a0[j] = __ffma2_rn(gj, __fadd2_rn(xv, nmu[j]), a0[j]);
            a1[j] = __fadd2_rn(a1[j], gj);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][threadIdx.x][2 * j] = a0[j].x, red[0][threadIdx.x][2 * j + 1] = a0[j].y;
      red[1][threadIdx.x][2 * j] = a1[j].x, red[1][threadIdx.x][2 * j + 1] = a1[j].y;
    }
    __syncthreads();
    if (G.r_in == 0 && g < G.groups) {
      float s0[8], s1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s0[j] = 0.f, s1[j] = 0.f;
        for (int r = 0; r < G.rl; ++r)
          s0[j] += red[0][r * G.cg + G.g_in][j], s1[j] += red[1][r * G.cg + G.g_in][j];
        s0[j] *= invstd[g * 8 + j];
      }
      // the L2 atomic units, not HBM, bound this kernel for wide layers (blocks x 2C atomics):
      // 16-byte vector reductions cut the operation count 4x
      float* dg = dgamma + g * 8;
      float* db = dbeta + g * 8;
      if (((reinterpret_cast<uintptr_t>(dg) | reinterpret_cast<uintptr_t>(db)) & 15) == 0) {
        red_add_f32x4(dg, s0[0], s0[1], s0[2], s0[3]);
        red_add_f32x4(dg + 4, s0[4], s0[5], s0[6], s0[7]);
        red_add_f32x4(db, s1[0], s1[1], s1[2], s1[3]);
        red_add_f32x4(db + 4, s1[4], s1[5], s1[6], s1[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dg + j, s0[j]), atomicAdd(db + j, s1[j]);
      }
    }
    __syncthreads();
  }
}

// dx = A*g + B*x + K with per-channel A = gamma*invstd, B = -A*invstd*dgamma/M,
// K = -A*dbeta/M - B*mean.  Optionally also stores the masked g (gradient of the residual
// branch); dres may alias dy (same element read before it is written by the same thread).
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_apply_kernel(const __nv_bfloat16* dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    float* dgamma, float* dbeta,
                    const float* __restrict__ fscale, const float* __restrict__ fshift,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* dres, long long P, int C,
                    int relu, float inv_count, const int rev,
                    const float* __restrict__ sum_g, const float* __restrict__ sum_gx) {
  const Geo G = geo(C);
  if (G.r_in >= G.rl) return;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int g0 = blockIdx.y * G.cg; g0 < G.groups; g0 += gridDim.y * G.cg) {
    const int g = g0 + G.g_in;
    if (g >= G.groups) continue;
    const int c = g * 8;
    float2 cA[4], cB[4], cK[4], fs[4], fh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a[2], b[2], kk[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ch = c + 2 * j + h;
        const float is = invstd[ch];
        float dg, db;
        if (sum_g != nullptr) {
          // the reduction was fused into the producing data gradient's epilogue (igemm.cu):
          // sum(g) and sum(g x) arrive raw; one thread per channel publishes dgamma / dbeta
          db = sum_g[ch];
          dg = is * (sum_gx[ch] - mean[ch] * db);
          if (blockIdx.x == 0 && G.r_in == 0) {
            dgamma[ch] = dg;
            dbeta[ch] = db;
          }
        } else {
          dg = dgamma[ch];
          db = dbeta[ch];
        }
        a[h] = gamma[ch] * is;
        b[h] = -a[h] * is * dg * inv_count;
        kk[h] = -a[h] * db * inv_count - b[h] * mean[ch];
      }
      cA[j] = make_float2(a[0], a[1]);
      cB[j] = make_float2(b[0], b[1]);
      cK[j] = make_float2(kk[0], kk[1]);
      fs[j] = relu == 2 ? make_float2(fscale[c + 2 * j], fscale[c + 2 * j + 1])
                        : make_float2(0.f, 0.f);
      fh[j] = relu == 2 ? make_float2(fshift[c + 2 * j], fshift[c + 2 * j + 1])
                        : make_float2(0.f, 0.f);
    }
    const long long stride = static_cast<long long>(gridDim.x) * G.rl;
    for (long long p = static_cast<long long>(blockIdx.x) * G.rl + G.r_in; p < P;
         p += kRows * stride) {
      uint4 gq[kRows], xq[kRows], yq[kRows];
      uint32_t mq[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long pk = p + k * stride;
        const bool in = pk < P;
        const long long o = (rev ? P - 1 - pk : pk) * C + c;
        gq[k] = in ? *reinterpret_cast<const uint4*>(dy + o) : zero;  // may alias dres: no .nc
        xq[k] = in ? ld_nc_v4(x + o) : zero;
        yq[k] = (in && relu == 1) ? ld_nc_v4(y + o) : zero;
        mq[k] = (in && relu == 3) ? __ldg(reinterpret_cast<const uint8_t*>(y) + (o >> 3)) : 0u;
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long pk = p + k * stride;
        if (pk >= P) break;
        const long long o = (rev ? P - 1 - pk : pk) * C + c;
        const uint32_t gw[4] = {gq[k].x, gq[k].y, gq[k].z, gq[k].w};
        const uint32_t xw[4] = {xq[k].x, xq[k].y, xq[k].z, xq[k].w};
        const uint32_t yw[4] = {yq[k].x, yq[k].y, yq[k].z, yq[k].w};
        uint32_t ow[4], rw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xv = up2(xw[j]);
          const float2 gj = masked_g(relu, gw[j], xv, yw[j], mq[k], j, fs[j], fh[j]);
          const float2 ov = __ffma2_rn(cA[j], gj, __ffma2_rn(cB[j], xv, cK[j]));
          ow[j] = pack_bf16x2(ov.x, ov.y);
          rw[j] = pack_bf16x2(gj.x, gj.y);
        }
        *reinterpret_cast<uint4*>(dx + o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (dres != nullptr)
          *reinterpret_cast<uint4*>(dres + o) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
      }
    }
  }
}

// Forward apply: y = act(x * scale[c] + shift[c] (+ residual)); act 0 none, 1 relu, 2 relu6.
// Same thread geometry as the backward kernels (a thread keeps one 8-channel group, so scale /
// shift are loaded once, and four rows are in flight); the arithmetic runs on the packed
// fp32x2 pipe.  mask (optional): one bit per element, set where the pre-activation is > 0.
template <bool FIN>
__global__ void __launch_bounds__(kThreads, 3)
bn_fwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ mask, long long P, int C,
                    int act, const BnFinalize fin, const int rev) {
  const Geo G = geo(C);
  if (G.r_in >= G.rl) return;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const bool has_res = residual != nullptr;
  for (int g0 = blockIdx.y * G.cg; g0 < G.groups; g0 += gridDim.y * G.cg) {
    const int g = g0 + G.g_in;
    if (g >= G.groups) continue;
    const int c = g * 8;
    float2 sc[4], sh[4];
    if (FIN) {
      const bool writer = blockIdx.x == 0 && G.r_in == 0;   // one thread per channel group
      float s8[8], h8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = c + j;
        const float m = fin.sum[ch] / fin.count;
        const float var = fmaxf(fin.sumsq[ch] / fin.count - m * m, 0.f);
        const float is = rsqrtf(var + fin.eps);
        s8[j] = fin.gamma[ch] * is;
        h8[j] = fin.beta[ch] - m * s8[j];
        if (writer) {
          fin.mean[ch] = m;
          fin.invstd[ch] = is;
          fin.scale[ch] = s8[j];
          fin.shift[ch] = h8[j];
          if (fin.running_mean != nullptr) {
            const float unbiased = fin.count > 1.f ? var * fin.count / (fin.count - 1.f) : var;
            fin.running_mean[ch] = (1.f - fin.momentum) * fin.running_mean[ch] + fin.momentum * m;
            fin.running_var[ch] =
                (1.f - fin.momentum) * fin.running_var[ch] + fin.momentum * unbiased;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[j] = make_float2(s8[2 * j], s8[2 * j + 1]);
        sh[j] = make_float2(h8[2 * j], h8[2 * j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[j] = make_float2(scale[c + 2 * j], scale[c + 2 * j + 1]);
        sh[j] = make_float2(shift[c + 2 * j], shift[c + 2 * j + 1]);
      }
    }
    const long long stride = static_cast<long long>(gridDim.x) * G.rl;
    for (long long p = static_cast<long long>(blockIdx.x) * G.rl + G.r_in; p < P;
         p += kRows * stride) {
      uint4 xq[kRows], rq[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long pk = p + k * stride;
        const bool in = pk < P;
        const long long o = (rev ? P - 1 - pk : pk) * C + c;
        xq[k] = in ? ld_nc_v4(x + o) : zero;
        rq[k] = (in && has_res) ? ld_nc_v4(residual + o) : zero;
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long pk = p + k * stride;
        if (pk >= P) break;
        const long long o = (rev ? P - 1 - pk : pk) * C + c;
        const uint32_t xw[4] = {xq[k].x, xq[k].y, xq[k].z, xq[k].w};
        const uint32_t rw[4] = {rq[k].x, rq[k].y, rq[k].z, rq[k].w};
        uint32_t ow[4], bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 v = __ffma2_rn(make_float2(__uint_as_float(xw[j] << 16),
                                            __uint_as_float(xw[j] & 0xffff0000u)),
                                sc[j], sh[j]);
          if (has_res)
            v = __fadd2_rn(v, make_float2(__uint_as_float(rw[j] << 16),
                                          __uint_as_float(rw[j] & 0xffff0000u)));
          bits |= (v.x > 0.f ? 1u : 0u) << (2 * j);
          bits |= (v.y > 0.f ? 1u : 0u) << (2 * j + 1);
          if (act >= 1) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f);
          if (act == 2) v.x = fminf(v.x, 6.f), v.y = fminf(v.y, 6.f);
          ow[j] = pack_bf16x2(v.x, v.y);
        }
        *reinterpret_cast<uint4*>(y + o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (mask != nullptr) mask[o >> 3] = static_cast<uint8_t>(bits);
      }
    }
  }
}

// L2 hand-over hint (see bn_set_row_reverse in ops.h): the next launches walk the rows from the
// END of the tensor - that is where the kernel that produced (or last read) it stopped, so the
// first ~L2-size worth of reads hit the 126 MB L2 instead of HBM.  Host state, read at launch
// time (and therefore baked into a captured CUDA graph).
int g_row_reverse = 0;

inline dim3 red_grid(long long P, int C, int waves) {
  const int groups = C >> 3;
  const int cg = groups < kThreads ? groups : kThreads;
  const int rl = kThreads / cg;
  long long bx = (P + rl - 1) / rl;
  const int gy = (groups + cg - 1) / cg;
  const long long cap = (148 * waves + gy - 1) / gy;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(gy));
}

// grid = resident blocks per SM (launch bounds: 2) x 148 SMs x waves; tunable for experiments
inline int env_waves(const char* name, int dflt) {
  const char* v = getenv(name);
  return v != nullptr && atoi(v) > 0 ? atoi(v) : dflt;
}

}  // namespace

void bn_set_row_reverse(int flag) { g_row_reverse = flag ? 1 : 0; }

cudaError_t bn_apply(const void* x, const void* residual, const float* scale, const float* shift,
                     void* y, uint8_t* mask, long long P, int C, int act, cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_FWD", 3);
  bn_fwd_apply_kernel<false><<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual), scale,
      shift, static_cast<__nv_bfloat16*>(y), mask, P, C, act, BnFinalize(), g_row_reverse);
  return cudaGetLastError();
}

cudaError_t bn_apply_finalize(const void* x, const void* residual, void* y, uint8_t* mask,
                              long long P, int C, int act, const BnFinalize& fin, cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_FWD", 3);
  bn_fwd_apply_kernel<true><<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(residual), nullptr,
      nullptr, static_cast<__nv_bfloat16*>(y), mask, P, C, act, fin, g_row_reverse);
  return cudaGetLastError();
}

cudaError_t bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean,
                          const float* invstd, const float* fscale, const float* fshift,
                          long long P, int C, int relu, float* dgamma, float* dbeta,
                          cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_REDUCE", 2);
  bn_bwd_reduce_kernel<<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(y), mean, invstd, fscale, fshift, P, C, relu, dgamma,
      dbeta, g_row_reverse);
  return cudaGetLastError();
}

cudaError_t bn_bwd_apply(const void* dy, const void* x, const void* y, const float* gamma,
                         const float* mean, const float* invstd, float* dgamma,
                         float* dbeta, const float* fscale, const float* fshift, void* dx,
                         void* dres, long long P, int C, int relu, const float* sum_g,
                         const float* sum_gx, cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_APPLY", 2);
  bn_bwd_apply_kernel<<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(y), gamma, mean, invstd, dgamma, dbeta, fscale, fshift,
      static_cast<__nv_bfloat16*>(dx), static_cast<__nv_bfloat16*>(dres), P, C, relu,
      1.f / static_cast<float>(P), g_row_reverse, sum_g, sum_gx);
  return cudaGetLastError();
}

}  // namespace tfos
