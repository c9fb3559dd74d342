// Memory-bound layer kernels (NHWC bf16 activations, fp32 statistics):
// batch-norm statistics / apply / backward, ReLU, residual add, max / average
// pooling, softmax cross-entropy (vector and per-pixel), uint8 input decode.
// These are the non-GEMM ops the reference reaches through TF/cuDNN for its
// example models; every kernel moves 16 bytes per
// thread per access and fuses what the producing/consuming op allows.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ops.h"
#include "ptx.cuh"

namespace tfos {

namespace {

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z),
               d = unpack_bf16x2(v.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = v;
}

// Column reduction skeleton: a block owns `cg` channel groups (8 channels each)
// and 256/cg row lanes; rows are strided over the grid; the row lanes are then
// folded through shared memory and the block issues one atomic per channel.
constexpr int kRedThreads = 256;

template <int NQ, typename F>
__device__ __forceinline__ void column_reduce(long long P, int C, F body, float* const (&out)[NQ]) {
  const int groups = C >> 3;
  const int cg = groups < kRedThreads ? groups : kRedThreads;  // channel groups per block pass
  const int rl = kRedThreads / cg;                             // row lanes
  const int g_in = threadIdx.x % cg;
  const int r_in = threadIdx.x / cg;
  __shared__ float red[NQ][kRedThreads][8];
  for (int g0 = blockIdx.y * cg; g0 < groups; g0 += gridDim.y * cg) {
    const int g = g0 + g_in;
    float acc[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
    if (g < groups && r_in < rl) {
      for (long long p = static_cast<long long>(blockIdx.x) * rl + r_in; p < P;
           p += static_cast<long long>(gridDim.x) * rl)
        body(p, g * 8, acc);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) red[q][threadIdx.x][j] = acc[q][j];
    __syncthreads();
    if (r_in == 0 && g < groups) {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float s = 0.f;
          for (int r = 0; r < rl; ++r) s += red[q][r * cg + g_in][j];
          atomicAdd(out[q] + g * 8 + j, s);
        }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kRedThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, float* sum, float* sumsq) {
  float* const outs[2] = {sum, sumsq};
  column_reduce<2>(
      P, C,
      [&](long long p, int c, float(&acc)[2][8]) {
        float f[8];
        load8(x + p * C + c, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += f[j];
          acc[1][j] += f[j] * f[j];
        }
      },
      outs);
}

// sums -> (mean, invstd, scale, shift), running statistics; clears the sums so
// the next step's fused-epilogue atomics start from zero.
__global__ void bn_finalize_kernel(float* sum, float* sumsq, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float* mean,
                                   float* invstd, float* scale, float* shift, int C, float count,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = sum[c] / count;
  const float var = fmaxf(sumsq[c] / count - m * m, 0.f);
  const float is = rsqrtf(var + eps);
  mean[c] = m;
  invstd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - m * sc;
  if (running_mean != nullptr) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
  sum[c] = 0.f;
  sumsq[c] = 0.f;
}

// inference-mode scale/shift from running statistics
__global__ void bn_inference_coeffs_kernel(const float* gamma, const float* beta,
                                           const float* running_mean, const float* running_var,
                                           float* scale, float* shift, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] * rsqrtf(running_var[c] + eps);
  scale[c] = sc;
  shift[c] = beta[c] - running_mean[c] * sc;
}

// (batch-norm apply and backward live in bn_bwd.cu)

// generic fused elementwise: out = act(a (+ b)); and relu backward
__global__ void __launch_bounds__(256)
add_act_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
               __nv_bfloat16* __restrict__ out, long long total8, int act) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], r[8];
    load8(a + i * 8, f);
    if (b != nullptr) {
      load8(b + i * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (act >= 1) f[j] = fmaxf(f[j], 0.f);
      if (act == 2) f[j] = fminf(f[j], 6.f);
    }
    store8(out + i * 8, f);
  }
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                __nv_bfloat16* __restrict__ dx, long long total8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float g[8], yv[8];
    load8(dy + i * 8, g);
    load8(y + i * 8, yv);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = yv[j] > 0.f ? g[j] : 0.f;
    store8(dx + i * 8, g);
  }
}

// column sums of a bf16 matrix into fp32 (bias gradients)
__global__ void __launch_bounds__(kRedThreads)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long P, int C, float* out) {
  float* const outs[1] = {out};
  column_reduce<1>(
      P, C,
      [&](long long p, int c, float(&acc)[1][8]) {
        float f[8];
        load8(x + p * C + c, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0][j] += f[j];
      },
      outs);
}

// ----------------------------------------------------------------- pooling
// max pool k x k, stride s, pad p over NHWC; records the argmax tap so the
// backward routes the gradient to exactly one input (PyTorch semantics).
// K > 0: window known at compile time - the K*K 16-byte loads of a thread are all in flight
// at once (ResNet's 3x3/2 pool); K == 0: run-time window.
template <int K, int S, int PAD>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                   uint8_t* __restrict__ idx, int N, int H, int W, int C, int OH, int OW, int k_rt,
                   int s_rt, int pad_rt) {
  const int k = K > 0 ? K : k_rt, s = K > 0 ? S : s_rt, pad = K > 0 ? PAD : pad_rt;
  const int groups = C >> 3;
  const long long total = static_cast<long long>(N) * OH * OW * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long long r = i / groups;
    const int ow = static_cast<int>(r % OW);
    r /= OW;
    const int oh = static_cast<int>(r % OH);
    const int n = static_cast<int>(r / OH);
    float best[8];
    int bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY, bi[j] = 0;
    if (K > 0) {
      uint4 raw[K > 0 ? K * K : 1];
      bool ok[K > 0 ? K * K : 1];
#pragma unroll
      for (int t = 0; t < K * K; ++t) {
        const int h = oh * S - PAD + t / (K > 0 ? K : 1), w = ow * S - PAD + t % (K > 0 ? K : 1);
        ok[t] = h >= 0 && h < H && w >= 0 && w < W;
        raw[t] = make_uint4(0u, 0u, 0u, 0u);
        if (ok[t])
          raw[t] = ld_nc_v4(x + ((static_cast<long long>(n) * H + h) * W + w) * C + g * 8);
      }
#pragma unroll
      for (int t = 0; t < K * K; ++t) {
        if (!ok[t]) continue;
        const uint32_t wd[4] = {raw[t].x, raw[t].y, raw[t].z, raw[t].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f = __uint_as_float((j & 1) ? (wd[j >> 1] & 0xffff0000u) : (wd[j >> 1] << 16));
          if (f > best[j]) best[j] = f, bi[j] = t;
        }
      }
    } else {
      for (int kh = 0; kh < k; ++kh) {
        const int h = oh * s - pad + kh;
        if (h < 0 || h >= H) continue;
        for (int kw = 0; kw < k; ++kw) {
          const int w = ow * s - pad + kw;
          if (w < 0 || w >= W) continue;
          float f[8];
          load8(x + ((static_cast<long long>(n) * H + h) * W + w) * C + g * 8, f);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (f[j] > best[j]) best[j] = f[j], bi[j] = kh * k + kw;
        }
      }
    }
    store8(y + i * 8, best);
    if (idx != nullptr) {
      uint2 p;
      p.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      p.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      *reinterpret_cast<uint2*>(idx + i * 8) = p;
    }
  }
}

// 3x3 / stride 2 / pad 1 with a 2-D thread tile: a block covers 8 x 4 output pixels x 8 channel
// groups (C = 64: one block per tile; wider C: blockIdx.y walks the channel groups), so the 3x3
// windows of neighbouring outputs - which share 5 of 9 input pixels - hit in L1 instead of
// pulling every input pixel 2.25 times through the L2 -> SM fabric (row-major thread order put
// the vertical neighbours in different blocks).
__global__ void __launch_bounds__(256)
maxpool_fwd_3x3s2_tiled_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                               uint8_t* __restrict__ idx, int N, int H, int W, int C, int OH,
                               int OW) {
  const int tiles_w = (OW + 7) >> 3, tiles_h = (OH + 3) >> 2;
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int n = t / tiles_h;
  const int g = blockIdx.y * 8 + (threadIdx.x & 7);
  const int ow = tw * 8 + ((threadIdx.x >> 3) & 7);
  const int oh = th * 4 + (threadIdx.x >> 6);
  if (ow >= OW || oh >= OH || g * 8 >= C) return;
  float best[8];
  int bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) best[j] = -INFINITY, bi[j] = 0;
  uint4 raw[9];
  bool ok[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int h = oh * 2 - 1 + k / 3, w = ow * 2 - 1 + k % 3;
    ok[k] = h >= 0 && h < H && w >= 0 && w < W;
    raw[k] = make_uint4(0u, 0u, 0u, 0u);
    if (ok[k]) raw[k] = ld_nc_v4(x + ((static_cast<long long>(n) * H + h) * W + w) * C + g * 8);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (!ok[k]) continue;
    const uint32_t wd[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = __uint_as_float((j & 1) ? (wd[j >> 1] & 0xffff0000u) : (wd[j >> 1] << 16));
      if (f > best[j]) best[j] = f, bi[j] = k;
    }
  }
  const long long o = ((static_cast<long long>(n) * OH + oh) * OW + ow) * C + g * 8;
  store8(y + o, best);
  if (idx != nullptr) {
    uint2 p;
    p.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    p.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + o) = p;
  }
}

__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int OH, int OW,
                   int k, int s, int pad) {
  const int groups = C >> 3;
  const long long total = static_cast<long long>(N) * H * W * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long long r = i / groups;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // outputs whose window covers (h, w): oh*s - pad <= h < oh*s - pad + k
    const int oh_lo = max(0, (h + pad - k + s) / s), oh_hi = min(OH - 1, (h + pad) / s);
    const int ow_lo = max(0, (w + pad - k + s) / s), ow_hi = min(OW - 1, (w + pad) / s);
    for (int oh = oh_lo; oh <= oh_hi; ++oh)
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int tap = (h - (oh * s - pad)) * k + (w - (ow * s - pad));
        const long long o = ((static_cast<long long>(n) * OH + oh) * OW + ow) * C + g * 8;
        float gv[8];
        load8(dy + o, gv);
        const uint2 p = *reinterpret_cast<const uint2*>(idx + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = ((j < 4 ? p.x : p.y) >> ((j & 3) * 8)) & 0xff;
          if (t == tap) acc[j] += gv[j];
        }
      }
    store8(dx + i * 8, acc);
  }
}

// 3x3 / stride 2 / pad 1: the 2x2 input block (2m..2m+1, 2n..2n+1) is covered by the four
// windows (m..m+1, n..n+1) only - four (dy, idx) loads produce four dx pixels.
__global__ void __launch_bounds__(256)
maxpool_bwd_3x3s2_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                         __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int OH,
                         int OW) {
  const int groups = C >> 3;
  const int BH = (H + 1) >> 1, BW = (W + 1) >> 1;
  // thread order: channel group, then an 8 x 4 tile of 2x2 blocks, then tiles - neighbouring
  // blocks (which share their pooling windows) sit in the same thread block and hit in L1
  const int tiles_w = (BW + 7) >> 3, tiles_h = (BH + 3) >> 2;
  const long long total = static_cast<long long>(N) * tiles_h * tiles_w * 32 * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long long r = i / groups;
    const int pw = static_cast<int>(r & 7);
    const int ph = static_cast<int>((r >> 3) & 3);
    r >>= 5;
    const int bw = static_cast<int>(r % tiles_w) * 8 + pw;
    r /= tiles_w;
    const int bh = static_cast<int>(r % tiles_h) * 4 + ph;
    const int n = static_cast<int>(r / tiles_h);
    if (bw >= BW || bh >= BH) continue;
    uint4 gq[4];
    uint2 iq[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int oh = bh + (t >> 1), ow = bw + (t & 1);
      gq[t] = make_uint4(0u, 0u, 0u, 0u);
      iq[t] = make_uint2(0xffffffffu, 0xffffffffu);  // tap 255 never matches
      if (oh < OH && ow < OW) {
        const long long o = ((static_cast<long long>(n) * OH + oh) * OW + ow) * C + g * 8;
        gq[t] = ld_nc_v4(dy + o);
        iq[t] = *reinterpret_cast<const uint2*>(idx + o);
      }
    }
    // tap (kh*3+kw) through which window t sees pixel (dh, dw) of the block; -1: not covered
    //   window (m,n):   (0,0)->4 (0,1)->5 (1,0)->7 (1,1)->8
    //   window (m,n+1): (0,1)->3 (1,1)->6      window (m+1,n): (1,0)->1 (1,1)->2
    //   window (m+1,n+1): (1,1)->0
    const int tapmap[4][4] = {{4, -1, -1, -1}, {5, 3, -1, -1}, {7, -1, 1, -1}, {8, 6, 2, 0}};
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const int h = 2 * bh + (px >> 1), w = 2 * bw + (px & 1);
      if (h >= H || w >= W) continue;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int tap = tapmap[px][t];
        if (tap < 0) continue;
        const uint32_t wd[4] = {gq[t].x, gq[t].y, gq[t].z, gq[t].w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int tj = ((j < 4 ? iq[t].x : iq[t].y) >> ((j & 3) * 8)) & 0xff;
          const float f = __uint_as_float((j & 1) ? (wd[j >> 1] & 0xffff0000u) : (wd[j >> 1] << 16));
          if (tj == tap) acc[j] += f;
        }
      }
      store8(dx + ((static_cast<long long>(n) * H + h) * W + w) * C + g * 8, acc);
    }
  }
}

// global average pool [N, HW, C] -> [N, C] and its backward
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                   int HW, int C) {
  const int groups = C >> 3;
  const long long total = static_cast<long long>(N) * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long n = i / groups;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int p = 0; p < HW; ++p) {
      float f[8];
      load8(x + (n * HW + p) * C + g * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    const float inv = 1.f / HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    store8(y + i * 8, acc);
  }
}

__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N,
                   int HW, int C) {
  const int groups = C >> 3;
  const long long total = static_cast<long long>(N) * HW * groups;
  const float inv = 1.f / HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    const long long n = i / groups / HW;
    float f[8];
    load8(dy + n * C + g * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    store8(dx + i * 8, f);
  }
}

// -------------------------------------------------------------------- loss
// One block per row: loss_sum += -log softmax(logits)[label] * scale and
// dlogits = (softmax - onehot) * scale.  Rows = samples (classification) or
// pixels (segmentation); ld = row pitch of logits/dlogits in elements.
template <typename T>
__device__ __forceinline__ float ld_logit(const T* p);
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

template <typename T>
__global__ void __launch_bounds__(128)
softmax_xent_kernel(const T* __restrict__ logits, const int* __restrict__ labels,
                    __nv_bfloat16* __restrict__ dlogits, float* loss_sum, float* correct_sum,
                    int V, int ld, int ldd, float scale) {
  const long long row = blockIdx.x;
  const T* l = logits + row * ld;
  __shared__ float red[4];
  __shared__ int redi[4];
  float mx = -INFINITY;
  int amax = 0;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    const float v = ld_logit(l + j);
    if (v > mx) mx = v, amax = j;
  }
  for (int o = 16; o >= 1; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffff, mx, o);
    const int oi = __shfl_xor_sync(0xffffffff, amax, o);
    if (ov > mx || (ov == mx && oi < amax)) mx = ov, amax = oi;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx, redi[threadIdx.x >> 5] = amax;
  __syncthreads();
  mx = red[0], amax = redi[0];
  for (int w = 1; w < 4; ++w)
    if (red[w] > mx || (red[w] == mx && redi[w] < amax)) mx = red[w], amax = redi[w];
  __syncthreads();
  float se = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) se += __expf(ld_logit(l + j) - mx);
  for (int o = 16; o >= 1; o >>= 1) se += __shfl_xor_sync(0xffffffff, se, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  se = red[0] + red[1] + red[2] + red[3];
  const int label = labels[row];
  const float inv = 1.f / se;
  if (dlogits != nullptr) {
    __nv_bfloat16* d = dlogits + row * ldd;
    for (int j = threadIdx.x; j < ldd; j += blockDim.x) {
      float g = 0.f;
      if (j < V) g = (__expf(ld_logit(l + j) - mx) * inv - (j == label ? 1.f : 0.f)) * scale;
      d[j] = __float2bfloat16_rn(g);
    }
  }
  if (threadIdx.x == 0) {
    const float lp = ld_logit(l + label) - mx - __logf(se);
    atomicAdd(loss_sum, -lp * scale);
    if (correct_sum != nullptr && amax == label) atomicAdd(correct_sum, 1.f);
  }
}

// -------------------------------------------------------------- input decode
// uint8 NHWC image -> normalised bf16 with Cp (>= C) channels and an optional
// zero border in W (the stem convolution's TMA window layout wants one).
__global__ void __launch_bounds__(256)
decode_normalize_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, int N,
                        int H, int W, int C, int Wp, int Cp, int wofs, float m0, float m1,
                        float m2, float s0, float s1, float s2) {
  const long long total = static_cast<long long>(N) * H * Wp;
  const float mean[3] = {m0, m1, m2};
  const float istd[3] = {s0, s1, s2};
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(i % Wp);
    const long long nh = i / Wp;
    const int w = wp - wofs;
    __nv_bfloat16* o = out + i * Cp;
    const bool inside = w >= 0 && w < W;
    const uint8_t* p = in + (nh * W + (inside ? w : 0)) * C;
    if ((Cp & 7) == 0) {
      for (int c0 = 0; c0 < Cp; c0 += 8) {  // one 16-byte store per 8 channels
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + j;
          v[j] = (inside && c < C)
                     ? (static_cast<float>(p[c]) * (1.f / 255.f) - mean[c % 3]) * istd[c % 3]
                     : 0.f;
        }
        store8(o + c0, v);
      }
    } else {
      for (int c = 0; c < Cp; ++c) {
        float v = 0.f;
        if (inside && c < C)
          v = (static_cast<float>(p[c]) * (1.f / 255.f) - mean[c % 3]) * istd[c % 3];
        o[c] = __float2bfloat16_rn(v);
      }
    }
  }
}

// RGB -> 8 padded channels (the stem layout): four consecutive padded pixels per thread, no
// per-channel loops or dynamic indexing - four independent 16-byte stores in flight.
__global__ void __launch_bounds__(256)
decode_rgb8_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, long long rows,
                   int W, int Wp, int wofs, float m0, float m1, float m2, float s0, float s1,
                   float s2) {
  const int quads = (Wp + 3) >> 2;
  const long long total = rows * quads;
  const float a0 = s0 * (1.f / 255.f), a1 = s1 * (1.f / 255.f), a2 = s2 * (1.f / 255.f);
  const float b0 = -m0 * s0, b1 = -m1 * s1, b2 = -m2 * s2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int qd = static_cast<int>(i % quads);
    const long long row = i / quads;
    const uint8_t* src = in + row * W * 3;
    __nv_bfloat16* dst = out + (row * Wp + qd * 4) * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int wp = qd * 4 + k;
      if (wp >= Wp) break;
      const int w = wp - wofs;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (w >= 0 && w < W) {
        const uint8_t* p = src + w * 3;
        const float r = static_cast<float>(p[0]) * a0 + b0;
        const float g = static_cast<float>(p[1]) * a1 + b1;
        const float b = static_cast<float>(p[2]) * a2 + b2;
        v.x = pack_bf16x2(r, g);
        v.y = pack_bf16x2(b, 0.f);
      }
      *reinterpret_cast<uint4*>(dst + k * 8) = v;
    }
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}

// Strided channel copy: dst[p, dst_off + c] = src[p, src_off + c], c < C (8-channel vectors).
// Writes an activation into its slot of a channel-concatenated buffer and cuts the matching
// slice out of the concatenated gradient on the way back (U-Net skip connections).
__global__ void __launch_bounds__(256)
copy_channels_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                     long long P, int C, int src_ld, int src_off, int dst_ld, int dst_off) {
  const int groups = C >> 3;
  const long long total = P * groups;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / groups;
    const int g = static_cast<int>(i % groups);
    *reinterpret_cast<uint4*>(dst + p * dst_ld + dst_off + g * 8) =
        *reinterpret_cast<const uint4*>(src + p * src_ld + src_off + g * 8);
  }
}

// Per-pixel softmax cross-entropy for a handful of classes (V <= 8): one thread per pixel, the
// pixel's 8 padded logits are one 16-byte load.  loss_sum += -log p[label] * scale,
// dlogits = (p - onehot) * scale, padded channels get zero gradient.
__global__ void __launch_bounds__(256)
pixel_xent_kernel(const __nv_bfloat16* __restrict__ logits, const int* __restrict__ labels,
                  __nv_bfloat16* __restrict__ dlogits, float* loss_sum, float* correct_sum,
                  long long P, int V, float scale) {
  float local = 0.f, hits = 0.f;
  for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < P;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    float l[8];
    load8(logits + p * 8, l);
    float mx = l[0];
    int am = 0;
#pragma unroll
    for (int j = 1; j < 8; ++j)
      if (j < V && l[j] > mx) mx = l[j], am = j;
    float se = 0.f, e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      e[j] = j < V ? __expf(l[j] - mx) : 0.f;
      se += e[j];
    }
    const int label = labels[p];
    const float inv = 1.f / se;
    float lab_logit = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j == label) lab_logit = l[j];
      e[j] = j < V ? (e[j] * inv - (j == label ? 1.f : 0.f)) * scale : 0.f;
    }
    if (dlogits != nullptr) store8(dlogits + p * 8, e);
    local += -(lab_logit - mx - __logf(se)) * scale;
    hits += (am == label) ? 1.f : 0.f;
  }
  for (int o = 16; o >= 1; o >>= 1) {
    local += __shfl_xor_sync(0xffffffff, local, o);
    hits += __shfl_xor_sync(0xffffffff, hits, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(loss_sum, local);
    if (correct_sum != nullptr) atomicAdd(correct_sum, hits);
  }
}

inline int grid_for(long long work, int threads, int max_blocks) {
  long long b = (work + threads - 1) / threads;
  if (b < 1) b = 1;
  return static_cast<int>(b < max_blocks ? b : max_blocks);
}
constexpr int kMaxBlocks = 148 * 8;

inline dim3 red_grid(long long P, int C, int waves = 4) {
  const int groups = C >> 3;
  const int cg = groups < kRedThreads ? groups : kRedThreads;
  const int rl = kRedThreads / cg;
  long long bx = (P + rl - 1) / rl;
  const int gy = (groups + cg - 1) / cg;
  const long long cap = (148 * waves + gy - 1) / gy;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3(static_cast<unsigned>(bx), static_cast<unsigned>(gy));
}

}  // namespace

#define TFOS_RET() return cudaGetLastError()

cudaError_t bn_stats(const void* x, long long P, int C, float* sum, float* sumsq, cudaStream_t s) {
  bn_stats_kernel<<<red_grid(P, C), kRedThreads, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), P, C, sum, sumsq);
  TFOS_RET();
}
cudaError_t bn_finalize(float* sum, float* sumsq, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, float* mean, float* invstd,
                        float* scale, float* shift, int C, float count, float eps, float momentum,
                        cudaStream_t s) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(sum, sumsq, gamma, beta, running_mean,
                                                     running_var, mean, invstd, scale, shift, C,
                                                     count, eps, momentum);
  TFOS_RET();
}
cudaError_t bn_inference_coeffs(const float* gamma, const float* beta, const float* rm,
                                const float* rv, float* scale, float* shift, int C, float eps,
                                cudaStream_t s) {
  bn_inference_coeffs_kernel<<<(C + 127) / 128, 128, 0, s>>>(gamma, beta, rm, rv, scale, shift, C,
                                                             eps);
  TFOS_RET();
}
cudaError_t add_act(const void* a, const void* b, void* out, long long n, int act,
                    cudaStream_t s) {
  add_act_kernel<<<grid_for(n / 8, 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b),
      static_cast<__nv_bfloat16*>(out), n / 8, act);
  TFOS_RET();
}
cudaError_t relu_bwd(const void* dy, const void* y, void* dx, long long n, cudaStream_t s) {
  relu_bwd_kernel<<<grid_for(n / 8, 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(dx), n / 8);
  TFOS_RET();
}
cudaError_t colsum(const void* x, long long P, int C, float* out, cudaStream_t s) {
  colsum_kernel<<<red_grid(P, C), kRedThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), P, C,
                                                       out);
  TFOS_RET();
}
cudaError_t maxpool_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int OH,
                        int OW, int k, int stride, int pad, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * OH * OW * (C >> 3);
  if (k == 3 && stride == 2 && pad == 1) {
    const long long tiles = static_cast<long long>(N) * ((OH + 3) / 4) * ((OW + 7) / 8);
    const int gy = ((C >> 3) + 7) / 8;
    if (tiles <= 0x7fffffffLL && gy <= 65535) {
      maxpool_fwd_3x3s2_tiled_kernel<<<dim3(static_cast<unsigned>(tiles), gy), 256, 0, s>>>(
          static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, N, H, W, C,
          OH, OW);
      TFOS_RET();
    }
    maxpool_fwd_kernel<3, 2, 1><<<grid_for(total, 256, kMaxBlocks * 4), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, N, H, W, C, OH,
        OW, k, stride, pad);
  } else
    maxpool_fwd_kernel<0, 0, 0><<<grid_for(total, 256, kMaxBlocks * 4), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, N, H, W, C, OH,
        OW, k, stride, pad);
  TFOS_RET();
}
cudaError_t maxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                        int OH, int OW, int k, int stride, int pad, cudaStream_t s) {
  if (k == 3 && stride == 2 && pad == 1) {
    const long long blocks = static_cast<long long>(N) * (((H + 1) / 2 + 3) / 4) *
                             (((W + 1) / 2 + 7) / 8) * 32 * (C >> 3);
    maxpool_bwd_3x3s2_kernel<<<grid_for(blocks, 256, kMaxBlocks * 4), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(dy), idx, static_cast<__nv_bfloat16*>(dx), N, H, W, C,
        OH, OW);
    TFOS_RET();
  }
  const long long total = static_cast<long long>(N) * H * W * (C >> 3);
  maxpool_bwd_kernel<<<grid_for(total, 256, kMaxBlocks * 4), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), idx, static_cast<__nv_bfloat16*>(dx), N, H, W, C, OH,
      OW, k, stride, pad);
  TFOS_RET();
}
cudaError_t avgpool_fwd(const void* x, void* y, int N, int HW, int C, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * (C >> 3);
  avgpool_fwd_kernel<<<grid_for(total, 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), N, HW, C);
  TFOS_RET();
}
cudaError_t avgpool_bwd(const void* dy, void* dx, int N, int HW, int C, cudaStream_t s) {
  const long long total = static_cast<long long>(N) * HW * (C >> 3);
  avgpool_bwd_kernel<<<grid_for(total, 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dx), N, HW, C);
  TFOS_RET();
}
cudaError_t softmax_xent(const void* logits, int logits_fp32, const int* labels, void* dlogits,
                         float* loss_sum, float* correct_sum, long long rows, int V, int ld,
                         int ldd, float scale, cudaStream_t s) {
  if (logits_fp32)
    softmax_xent_kernel<float><<<static_cast<unsigned>(rows), 128, 0, s>>>(
        static_cast<const float*>(logits), labels, static_cast<__nv_bfloat16*>(dlogits), loss_sum,
        correct_sum, V, ld, ldd, scale);
  else
    softmax_xent_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), 128, 0, s>>>(
        static_cast<const __nv_bfloat16*>(logits), labels, static_cast<__nv_bfloat16*>(dlogits),
        loss_sum, correct_sum, V, ld, ldd, scale);
  TFOS_RET();
}
cudaError_t decode_normalize(const uint8_t* in, void* out, int N, int H, int W, int C, int Wp,
                             int Cp, int wofs, const float* mean3, const float* istd3,
                             cudaStream_t s) {
  if (C == 3 && Cp == 8) {
    const long long rows = static_cast<long long>(N) * H;
    decode_rgb8_kernel<<<grid_for(rows * ((Wp + 3) / 4), 256, kMaxBlocks * 4), 256, 0, s>>>(
        in, static_cast<__nv_bfloat16*>(out), rows, W, Wp, wofs, mean3[0], mean3[1], mean3[2],
        istd3[0], istd3[1], istd3[2]);
    TFOS_RET();
  }
  const long long total = static_cast<long long>(N) * H * Wp;
  decode_normalize_kernel<<<grid_for(total, 256, kMaxBlocks * 4), 256, 0, s>>>(
      in, static_cast<__nv_bfloat16*>(out), N, H, W, C, Wp, Cp, wofs, mean3[0], mean3[1], mean3[2],
      istd3[0], istd3[1], istd3[2]);
  TFOS_RET();
}
cudaError_t copy_channels(const void* src, void* dst, long long P, int C, int src_ld, int src_off,
                          int dst_ld, int dst_off, cudaStream_t s) {
  copy_channels_kernel<<<grid_for(P * (C >> 3), 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), P, C, src_ld,
      src_off, dst_ld, dst_off);
  TFOS_RET();
}
cudaError_t pixel_xent(const void* logits, const int* labels, void* dlogits, float* loss_sum,
                       float* correct_sum, long long P, int V, float scale, cudaStream_t s) {
  pixel_xent_kernel<<<grid_for(P, 256, kMaxBlocks), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(logits), labels, static_cast<__nv_bfloat16*>(dlogits),
      loss_sum, correct_sum, P, V, scale);
  TFOS_RET();
}
cudaError_t cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t s) {
  cast_f32_bf16_kernel<<<grid_for(n, 256, kMaxBlocks), 256, 0, s>>>(
      in, static_cast<__nv_bfloat16*>(out), n);
  TFOS_RET();
}

}  // namespace tfos
