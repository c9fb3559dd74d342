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

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16), f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16), f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16), f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16), f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
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
                     float* dbeta) {
  const Geo G = geo(C);
  __shared__ float red[2][kThreads][8];
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int g0 = blockIdx.y * G.cg; g0 < G.groups; g0 += gridDim.y * G.cg) {
    const int g = g0 + G.g_in;
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = 0.f, a1[j] = 0.f;
    if (g < G.groups && G.r_in < G.rl) {
      const int c = g * 8;
      float mu[8], fs[8], fh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mu[j] = mean[c + j];
        fs[j] = relu == 2 ? fscale[c + j] : 0.f;
        fh[j] = relu == 2 ? fshift[c + j] : 0.f;
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
          const long long o = pk * C + c;
          gq[k] = in ? ld_nc_v4(dy + o) : zero;
          xq[k] = in ? ld_nc_v4(x + o) : zero;
          yq[k] = (in && relu == 1) ? ld_nc_v4(y + o) : zero;
          mq[k] = (in && relu == 3) ? __ldg(reinterpret_cast<const uint8_t*>(y) + (o >> 3)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          float gv[8], xv[8], yv[8];
          unpack8(gq[k], gv);
          unpack8(xq[k], xv);
          if (relu == 1) unpack8(yq[k], yv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float pre = relu == 2   ? xv[j] * fs[j] + fh[j]
                              : relu == 1 ? yv[j]
                              : relu == 3 ? (((mq[k] >> j) & 1u) ? 1.f : 0.f)
                                          : 1.f;
            const float gj = pre > 0.f ? gv[j] : 0.f;
            a0[j] += gj * (xv[j] - mu[j]);
            a1[j] += gj;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[0][threadIdx.x][j] = a0[j], red[1][threadIdx.x][j] = a1[j];
    __syncthreads();
    if (G.r_in == 0 && g < G.groups) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float s0 = 0.f, s1 = 0.f;
        for (int r = 0; r < G.rl; ++r)
          s0 += red[0][r * G.cg + G.g_in][j], s1 += red[1][r * G.cg + G.g_in][j];
        atomicAdd(dgamma + g * 8 + j, s0 * invstd[g * 8 + j]);
        atomicAdd(dbeta + g * 8 + j, s1);
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
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                    const float* __restrict__ fscale, const float* __restrict__ fshift,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* dres, long long P, int C,
                    int relu, float inv_count) {
  const Geo G = geo(C);
  if (G.r_in >= G.rl) return;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int g0 = blockIdx.y * G.cg; g0 < G.groups; g0 += gridDim.y * G.cg) {
    const int g = g0 + G.g_in;
    if (g >= G.groups) continue;
    const int c = g * 8;
    float cA[8], cB[8], cK[8], fs[8], fh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      fs[j] = relu == 2 ? fscale[c + j] : 0.f;
      fh[j] = relu == 2 ? fshift[c + j] : 0.f;
      const float is = invstd[c + j];
      const float a = gamma[c + j] * is;
      const float b = -a * is * dgamma[c + j] * inv_count;
      cA[j] = a;
      cB[j] = b;
      cK[j] = -a * dbeta[c + j] * inv_count - b * mean[c + j];
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
        const long long o = pk * C + c;
        gq[k] = in ? *reinterpret_cast<const uint4*>(dy + o) : zero;  // may alias dres: no .nc
        xq[k] = in ? ld_nc_v4(x + o) : zero;
        yq[k] = (in && relu == 1) ? ld_nc_v4(y + o) : zero;
        mq[k] = (in && relu == 3) ? __ldg(reinterpret_cast<const uint8_t*>(y) + (o >> 3)) : 0u;
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const long long pk = p + k * stride;
        if (pk >= P) break;
        const long long o = pk * C + c;
        float gv[8], xv[8], yv[8], ov[8];
        unpack8(gq[k], gv);
        unpack8(xq[k], xv);
        if (relu == 1) unpack8(yq[k], yv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pre = relu == 2   ? xv[j] * fs[j] + fh[j]
                            : relu == 1 ? yv[j]
                            : relu == 3 ? (((mq[k] >> j) & 1u) ? 1.f : 0.f)
                                        : 1.f;
          gv[j] = pre > 0.f ? gv[j] : 0.f;
          ov[j] = cA[j] * gv[j] + cB[j] * xv[j] + cK[j];
        }
        *reinterpret_cast<uint4*>(dx + o) = pack8(ov);
        if (dres != nullptr) *reinterpret_cast<uint4*>(dres + o) = pack8(gv);
      }
    }
  }
}

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

cudaError_t bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean,
                          const float* invstd, const float* fscale, const float* fshift,
                          long long P, int C, int relu, float* dgamma, float* dbeta,
                          cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_REDUCE", 2);
  bn_bwd_reduce_kernel<<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(y), mean, invstd, fscale, fshift, P, C, relu, dgamma,
      dbeta);
  return cudaGetLastError();
}

cudaError_t bn_bwd_apply(const void* dy, const void* x, const void* y, const float* gamma,
                         const float* mean, const float* invstd, const float* dgamma,
                         const float* dbeta, const float* fscale, const float* fshift, void* dx,
                         void* dres, long long P, int C, int relu, cudaStream_t s) {
  static const int waves = env_waves("TFOS_BN_WAVES_APPLY", 2);
  bn_bwd_apply_kernel<<<red_grid(P, C, waves), kThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x),
      static_cast<const __nv_bfloat16*>(y), gamma, mean, invstd, dgamma, dbeta, fscale, fshift,
      static_cast<__nv_bfloat16*>(dx), static_cast<__nv_bfloat16*>(dres), P, C, relu,
      1.f / static_cast<float>(P));
  return cudaGetLastError();
}

}  // namespace tfos
