// Fused collective + optimizer kernels over NVLink peer memory.
//
// The reference delegates the gradient all-reduce, the initial variable
// broadcast and the parameter-server push/pull to TensorFlow's runtime
// (MultiWorkerMirroredStrategy / ParameterServerStrategy; SURVEY.md section
// 2.6(a), reference call sites examples/mnist/keras/mnist_spark.py:11,55-66 and
// examples/mnist/estimator/mnist_spark_streaming.py:86).  Here each of them is
// ONE kernel that issues the peer loads/stores itself:
//
//   allreduce_opt : rank r owns shard r of a gradient bucket, pulls the 7 peer
//                   copies of that shard over NVLink (or one multimem.ld_reduce
//                   through the switch), averages in fp32, applies SGD /
//                   momentum / Adam to its fp32 master shard and stores the
//                   updated bf16 weights into every rank's weight buffer - the
//                   "all-gather" half of the all-reduce carries the new
//                   parameters instead of the reduced gradients.
//   bcast_pull    : startup broadcast, every rank pulls the root's parameters.
//   ps_*          : asynchronous parameter server: workers apply gradients to
//                   PS-resident weights with remote red.add (no barrier) and
//                   pull fresh parameters with peer loads.
//
// Cross-GPU ordering uses release/acquire at .sys scope on per-rank flag words
// living in the same symmetric allocation; every wait is bounded and traps.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ops.h"
#include "ptx.cuh"

namespace tfos {

namespace {

#ifndef TFOS_FLAG_TIMEOUT_NS
#define TFOS_FLAG_TIMEOUT_NS 20000000000ull
#endif
// how long a rank waits for a peer before it gives up (trap -> CUDA error -> node error queue ->
// driver exception, TFSparkNode.py): runtime knob, TFOS_FLAG_TIMEOUT_MS (parallel/symm.py)
__device__ unsigned long long g_flag_timeout_ns = TFOS_FLAG_TIMEOUT_NS;

__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t target) {
  if (static_cast<int32_t>(ld_acquire_sys(p) - target) >= 0) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (static_cast<int32_t>(ld_acquire_sys(p) - target) < 0) {
    if ((++spins & 0xff) == 0) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > g_flag_timeout_ns) {
        printf("tfos: peer flag timeout (block %d, want %u, have %u)\n", blockIdx.x, target,
               ld_acquire_sys(p));
        __trap();
      }
    }
  }
}

__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_b32x4(void* mc, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}

// hyper[]: 0 lr, 1 momentum, 2 weight_decay, 3 grad_scale (1/world * loss scale),
//          4 beta1, 5 beta2, 6 eps, 7 step (Adam bias correction)
template <int OPT>
__device__ __forceinline__ float opt_update(float w, float g, float& s1, float& s2,
                                            const float* h, bool decay) {
  g *= h[3];
  if (OPT == kOptAdam) {
    // decoupled-from-nothing: classic Adam with L2 folded into the gradient
    if (decay) g += h[2] * w;
    s1 = h[4] * s1 + (1.f - h[4]) * g;
    s2 = h[5] * s2 + (1.f - h[5]) * g * g;
    const float c1 = 1.f - __powf(h[4], h[7]);
    const float c2 = 1.f - __powf(h[5], h[7]);
    return w - h[0] * (s1 / c1) / (sqrtf(s2 / c2) + h[6]);
  }
  if (decay) g += h[2] * w;
  if (OPT == kOptMomentum) {
    s1 = h[1] * s1 + g;
    g = s1;
  }
  return w - h[0] * g;
}

// PHASE splits the kernel for the hierarchical (several hosts) all-reduce: 1 = intra-host
// reduce-scatter only (the sum over the local peers of this rank's shard replaces that shard of
// the rank's OWN gradient buffer - peers only ever read the other shards of it); 2 = update +
// all-gather only (the gradient comes from the own buffer, which by then holds the sum over all
// hosts; new weights go to every local peer).  0 = both in one pass (one host).
template <int OPT, bool MULTIMEM, int PHASE = 0>
__global__ void __launch_bounds__(512, 1) allreduce_opt_kernel(const AllreduceOptArgs a) {
  __shared__ float h[8];
  if (threadIdx.x < 8) h[threadIdx.x] = a.hyper[threadIdx.x];
  const int world = a.world, rank = a.rank;
  uint32_t e = 0;
  if (world > 1) {
    e = *a.epoch + 1;
    uint32_t* my = a.flags[rank] + a.slot * 32;
    if (blockIdx.x == 0 && threadIdx.x < world)
      st_release_sys(a.flags[threadIdx.x] + a.slot * 32 + rank, e);
    if (threadIdx.x < world) wait_flag(my + threadIdx.x, e);
  }
  __syncthreads();

  const long long n = a.end - a.begin;
  long long chunk = (n + world - 1) / world;
  chunk = (chunk + 7) & ~7ll;
  const long long lo = a.begin + min(n, chunk * rank);
  const long long hi = a.begin + min(n, chunk * (rank + 1));

  // NVLS: a thread keeps kU x 2 = 16 multimem.ld_reduce (16 bytes each, reduced in the switch) in
  // flight; the peer-to-peer path already has 2 x world loads per element group in flight
  constexpr int kU = MULTIMEM ? 8 : 1;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (long long i0 = lo + (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
       i0 < hi; i0 += stride * kU) {
   float4 mg0[kU], mg1[kU];
   if (MULTIMEM) {
#pragma unroll
     for (int u = 0; u < kU; ++u) {
       const long long iu = i0 + u * stride;
       if (iu < hi) {
         mg0[u] = multimem_ld_reduce_f32x4(a.grads_mc + iu);
         mg1[u] = multimem_ld_reduce_f32x4(a.grads_mc + iu + 4);
       }
     }
   }
#pragma unroll
   for (int u = 0; u < kU; ++u) {
    const long long i = i0 + u * stride;
    if (i >= hi) break;
    float g[8];
    if (world == 1 || PHASE == 2) {
      const float* own = (PHASE == 2) ? a.grads[rank] : a.grads[0];
      const float4 g0 = *reinterpret_cast<const float4*>(own + i);
      const float4 g1 = *reinterpret_cast<const float4*>(own + i + 4);
      g[0] = g0.x, g[1] = g0.y, g[2] = g0.z, g[3] = g0.w;
      g[4] = g1.x, g[5] = g1.y, g[6] = g1.z, g[7] = g1.w;
    } else if (MULTIMEM) {
      const float4 g0 = mg0[u], g1 = mg1[u];
      g[0] = g0.x, g[1] = g0.y, g[2] = g0.z, g[3] = g0.w;
      g[4] = g1.x, g[5] = g1.y, g[6] = g1.z, g[7] = g1.w;
    } else {
      uint4 v0[kMaxRanks], v1[kMaxRanks];
      // issue every peer load before the first use: 2*world 16-byte requests in flight
#pragma unroll
      for (int p = 0; p < kMaxRanks; ++p)
        if (p < world) {
          const int src = (rank + p) % world;  // stagger peers so links are used evenly
          v0[p] = ld_relaxed_sys_v4(a.grads[src] + i);
          v1[p] = ld_relaxed_sys_v4(a.grads[src] + i + 4);
        }
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = 0.f;
#pragma unroll
      for (int p = 0; p < kMaxRanks; ++p)
        if (p < world) {
          g[0] += __uint_as_float(v0[p].x), g[1] += __uint_as_float(v0[p].y);
          g[2] += __uint_as_float(v0[p].z), g[3] += __uint_as_float(v0[p].w);
          g[4] += __uint_as_float(v1[p].x), g[5] += __uint_as_float(v1[p].y);
          g[6] += __uint_as_float(v1[p].z), g[7] += __uint_as_float(v1[p].w);
        }
    }
    if (PHASE == 1) {
      *reinterpret_cast<float4*>(a.grads[rank] + i) = make_float4(g[0], g[1], g[2], g[3]);
      *reinterpret_cast<float4*>(a.grads[rank] + i + 4) = make_float4(g[4], g[5], g[6], g[7]);
      continue;
    }
    float w[8], s1[8], s2[8];
    const long long li = i - a.state_offset;  // master/state are indexed shard-locally
    {
      const float4 w0 = *reinterpret_cast<const float4*>(a.master + li);
      const float4 w1 = *reinterpret_cast<const float4*>(a.master + li + 4);
      w[0] = w0.x, w[1] = w0.y, w[2] = w0.z, w[3] = w0.w;
      w[4] = w1.x, w[5] = w1.y, w[6] = w1.z, w[7] = w1.w;
    }
    if (OPT != kOptSgd) {
      const float4 m0 = *reinterpret_cast<const float4*>(a.state1 + li);
      const float4 m1 = *reinterpret_cast<const float4*>(a.state1 + li + 4);
      s1[0] = m0.x, s1[1] = m0.y, s1[2] = m0.z, s1[3] = m0.w;
      s1[4] = m1.x, s1[5] = m1.y, s1[6] = m1.z, s1[7] = m1.w;
    }
    if (OPT == kOptAdam) {
      const float4 m0 = *reinterpret_cast<const float4*>(a.state2 + li);
      const float4 m1 = *reinterpret_cast<const float4*>(a.state2 + li + 4);
      s2[0] = m0.x, s2[1] = m0.y, s2[2] = m0.z, s2[3] = m0.w;
      s2[4] = m1.x, s2[5] = m1.y, s2[6] = m1.z, s2[7] = m1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = opt_update<OPT>(w[j], g[j], s1[j], s2[j], h, i + j < a.decay_end);
    *reinterpret_cast<float4*>(a.master + li) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(a.master + li + 4) = make_float4(w[4], w[5], w[6], w[7]);
    if (OPT != kOptSgd) {
      *reinterpret_cast<float4*>(a.state1 + li) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      *reinterpret_cast<float4*>(a.state1 + li + 4) = make_float4(s1[4], s1[5], s1[6], s1[7]);
    }
    if (OPT == kOptAdam) {
      *reinterpret_cast<float4*>(a.state2 + li) = make_float4(s2[0], s2[1], s2[2], s2[3]);
      *reinterpret_cast<float4*>(a.state2 + li + 4) = make_float4(s2[4], s2[5], s2[6], s2[7]);
    }
    uint4 packed;
    packed.x = pack_bf16x2(w[0], w[1]);
    packed.y = pack_bf16x2(w[2], w[3]);
    packed.z = pack_bf16x2(w[4], w[5]);
    packed.w = pack_bf16x2(w[6], w[7]);
    if (world == 1) {
      *reinterpret_cast<uint4*>(a.weights[0] + i) = packed;
    } else if (MULTIMEM) {
      multimem_st_b32x4(a.weights_mc + i, packed);
    } else {
#pragma unroll
      for (int p = 0; p < kMaxRanks; ++p)
        if (p < world) st_relaxed_sys_v4(a.weights[(rank + p) % world] + i, packed);
    }
    if (a.aux32[0] != nullptr && i >= a.aux_begin) {
      const long long ai = i - a.aux_begin;
      uint4 lo4, hi4;
      lo4.x = __float_as_uint(w[0]), lo4.y = __float_as_uint(w[1]);
      lo4.z = __float_as_uint(w[2]), lo4.w = __float_as_uint(w[3]);
      hi4.x = __float_as_uint(w[4]), hi4.y = __float_as_uint(w[5]);
      hi4.z = __float_as_uint(w[6]), hi4.w = __float_as_uint(w[7]);
      if (MULTIMEM && a.aux32_mc != nullptr) {
        multimem_st_b32x4(a.aux32_mc + ai, lo4);
        multimem_st_b32x4(a.aux32_mc + ai + 4, hi4);
      } else
#pragma unroll
      for (int p = 0; p < kMaxRanks; ++p)
        if (p < world) {
          float* dst = a.aux32[(rank + p) % world] + ai;
          if (world == 1) {
            *reinterpret_cast<uint4*>(dst) = lo4;
            *reinterpret_cast<uint4*>(dst + 4) = hi4;
          } else {
            st_relaxed_sys_v4(dst, lo4);
            st_relaxed_sys_v4(dst + 4, hi4);
          }
        }
    }
    if (a.zero_grads && world == 1) {
      *reinterpret_cast<float4*>(a.grads[0] + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(a.grads[0] + i + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
   }
  }

  if (world > 1) {
    // completion: the last block to finish publishes "rank done" to every peer
    // and returns only after every peer has published too, so kernel completion
    // on this rank implies its weight buffer holds all shards of the new weights
    // and no peer is still reading this rank's gradients.
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t is_last;
    if (threadIdx.x == 0) {
      const uint32_t prev = atomicAdd(a.block_counter, 1u);
      is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      if (threadIdx.x < world) {
        st_release_sys(a.flags[threadIdx.x] + a.slot * 32 + 16 + rank, e);
        wait_flag(a.flags[rank] + a.slot * 32 + 16 + threadIdx.x, e);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        *a.block_counter = 0;
        *a.epoch = e;
      }
    }
  }
}

// Every rank pulls [begin, end) of the root's buffers (bf16 weights and,
// optionally, an fp32 buffer) over NVLink.  Start/end flag barriers as above.
__global__ void __launch_bounds__(512, 1) bcast_pull_kernel(const BcastArgs a) {
  const int world = a.world, rank = a.rank;
  uint32_t e = 0;
  if (world > 1) {
    e = *a.epoch + 1;
    if (blockIdx.x == 0 && threadIdx.x < world)
      st_release_sys(a.flags[threadIdx.x] + a.slot * 32 + rank, e);
    if (threadIdx.x < world) wait_flag(a.flags[rank] + a.slot * 32 + threadIdx.x, e);
  }
  __syncthreads();
  if (rank != a.root) {
    const long long n16 = a.bytes / 16;
    const uint4* src = reinterpret_cast<const uint4*>(a.bufs[a.root]);
    uint4* dst = reinterpret_cast<uint4*>(a.bufs[rank]);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
         i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = i + static_cast<long long>(u) * gridDim.x * blockDim.x;
        if (k < n16) v[u] = ld_relaxed_sys_v4(src + k);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long k = i + static_cast<long long>(u) * gridDim.x * blockDim.x;
        if (k < n16) dst[k] = v[u];
      }
    }
  }
  if (world > 1) {
    __threadfence_system();
    __syncthreads();
    __shared__ uint32_t is_last;
    if (threadIdx.x == 0) is_last = (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (is_last) {
      if (threadIdx.x < world) {
        st_release_sys(a.flags[threadIdx.x] + a.slot * 32 + 16 + rank, e);
        wait_flag(a.flags[rank] + a.slot * 32 + 16 + threadIdx.x, e);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        *a.block_counter = 0;
        *a.epoch = e;
      }
    }
  }
}

// Standalone device-side barrier across ranks (metric sync / distributed save).
__global__ void flag_barrier_kernel(const BcastArgs a) {
  const uint32_t e = *a.epoch + 1;
  if (threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.slot * 32 + a.rank, e);
    wait_flag(a.flags[a.rank] + a.slot * 32 + threadIdx.x, e);
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.epoch = e;
}

// ------------------------------------------------------- parameter server
// Dense push: w_ps += -lr * scale * g applied with remote reductions; any number
// of workers may push concurrently (Hogwild-style asynchronous SGD).
__global__ void __launch_bounds__(256)
ps_push_dense_kernel(float* __restrict__ w_ps, const float* __restrict__ g, long long n,
                     const float* __restrict__ hyper) {
  const float k = -hyper[0] * hyper[3];
  // (a row cut by a server boundary arrives here with an arbitrary element offset)
  const bool vec = ((reinterpret_cast<uintptr_t>(w_ps) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    if (vec && i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(w_ps + i),
                   "f"(k * v.x), "f"(k * v.y), "f"(k * v.z), "f"(k * v.w)
                   : "memory");
    } else {
      for (long long j = i; j < min(n, i + 4); ++j)
        asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(w_ps + j), "f"(k * g[j])
                     : "memory");
    }
  }
}

// Sparse push (IndexedSlices): rows[idx[r]] += -lr * scale * grad_rows[r].  Rows whose width is
// a multiple of 4 (every embedding table in practice) go 16 bytes at a time: one index
// decomposition and one red.global.add.v4.f32 per four elements; duplicate indices simply add up.
__global__ void __launch_bounds__(256)
ps_push_sparse_kernel(float* __restrict__ w_ps, const float* __restrict__ g_rows,
                      const int* __restrict__ idx, int nrows, int width,
                      const float* __restrict__ hyper) {
  const float k = -hyper[0] * hyper[3];
  const bool vec = (width & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(w_ps) | reinterpret_cast<uintptr_t>(g_rows)) & 15) == 0;
  if (vec) {
    const int w4 = width >> 2;
    const long long total4 = static_cast<long long>(nrows) * w4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int r = static_cast<int>(i / w4), c = static_cast<int>(i - static_cast<long long>(r) * w4) * 4;
      const float4 v = *reinterpret_cast<const float4*>(g_rows + static_cast<long long>(r) * width + c);
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                       w_ps + static_cast<long long>(idx[r]) * width + c),
                   "f"(k * v.x), "f"(k * v.y), "f"(k * v.z), "f"(k * v.w)
                   : "memory");
    }
    return;
  }
  const long long total = static_cast<long long>(nrows) * width;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / width), c = static_cast<int>(i % width);
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(
                     w_ps + static_cast<long long>(idx[r]) * width + c),
                 "f"(k * g_rows[i])
                 : "memory");
  }
}

// Pull: read PS-resident fp32 parameters over NVLink, keep an fp32 copy and the
// bf16 compute copy locally (one pass).
__global__ void __launch_bounds__(256)
ps_pull_kernel(const float* __restrict__ w_ps, float* __restrict__ w_local,
               __nv_bfloat16* __restrict__ w_bf16, long long n) {
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    if (i + 4 <= n) {
      const uint4 v = ld_relaxed_sys_v4(w_ps + i);
      if (w_local != nullptr) *reinterpret_cast<uint4*>(w_local + i) = v;
      if (w_bf16 != nullptr) {
        uint2 p;
        p.x = pack_bf16x2(__uint_as_float(v.x), __uint_as_float(v.y));
        p.y = pack_bf16x2(__uint_as_float(v.z), __uint_as_float(v.w));
        *reinterpret_cast<uint2*>(w_bf16 + i) = p;
      }
    } else {
      for (long long j = i; j < n; ++j) {
        const float v = w_ps[j];
        if (w_local != nullptr) w_local[j] = v;
        if (w_bf16 != nullptr) w_bf16[j] = __float2bfloat16_rn(v);
      }
    }
  }
}

// ------------------------------------------- parameter server with optimizer state
// (parallel/ps.py "slot mode").  The PS GPU owns fp32 parameters + optimizer state + a bf16
// serving copy; every worker owns two gradient slots in the PS GPU's memory.
//   worker:  ps_push_slot  - waits (device side, bounded) until its slot has been applied, stores
//            its fp32 gradients into the slot with one-way NVLink writes, appends the batch-norm
//            running-statistics deltas, publishes a sequence number with st.release.sys;
//   server:  ps_apply      - applies SGD / momentum / Adam of ONE slot to the resident state at
//            HBM speed, refreshes the bf16 serving copy, marks the slot applied;
//   worker:  ps_pull_model - peer loads of the bf16 weights + the fp32 tail (BN scale/offset,
//            biases, running statistics).
// No barrier anywhere: workers never wait for each other (asynchronous SGD), only for their own
// slot (back-pressure when the server falls behind).
template <int OPT>
__global__ void __launch_bounds__(512) ps_apply_kernel(const PsApplyArgs a) {
  __shared__ float h[8];
  if (threadIdx.x < 8) h[threadIdx.x] = a.hyper[threadIdx.x];
  __syncthreads();
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < a.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
    const float4 g4 = *reinterpret_cast<const float4*>(a.slot + i);
    float4 w4 = *reinterpret_cast<const float4*>(a.master + i);
    float g[4] = {g4.x, g4.y, g4.z, g4.w};
    float w[4] = {w4.x, w4.y, w4.z, w4.w};
    const long long gi = a.lo + i;           // global element index of this server's slice
    if (gi >= a.ema_begin) {
      // non-trainable tail: the slot carries (pulled - locally updated) running statistics
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] -= g[j];
    } else {
      float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
      if (OPT != kOptSgd) {
        const float4 m = *reinterpret_cast<const float4*>(a.state1 + i);
        s1[0] = m.x, s1[1] = m.y, s1[2] = m.z, s1[3] = m.w;
      }
      if (OPT == kOptAdam) {
        const float4 v = *reinterpret_cast<const float4*>(a.state2 + i);
        s2[0] = v.x, s2[1] = v.y, s2[2] = v.z, s2[3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        w[j] = opt_update<OPT>(w[j], g[j], s1[j], s2[j], h, gi + j < a.decay_end);
      if (OPT != kOptSgd)
        *reinterpret_cast<float4*>(a.state1 + i) = make_float4(s1[0], s1[1], s1[2], s1[3]);
      if (OPT == kOptAdam)
        *reinterpret_cast<float4*>(a.state2 + i) = make_float4(s2[0], s2[1], s2[2], s2[3]);
    }
    *reinterpret_cast<float4*>(a.master + i) = make_float4(w[0], w[1], w[2], w[3]);
    uint2 pk;
    pk.x = pack_bf16x2(w[0], w[1]);
    pk.y = pack_bf16x2(w[2], w[3]);
    *reinterpret_cast<uint2*>(a.wbf16 + i) = pk;
  }
  // the last block publishes "slot applied" (workers poll it over NVLink before reusing the slot)
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    *a.block_counter = 0;
    st_release_sys(a.applied_flag, a.seq);
  }
}

__global__ void __launch_bounds__(512) ps_push_slot_kernel(const PsPushArgs a) {
  // back-pressure: the slot is free once the server has applied what this worker last put there
  if (threadIdx.x == 0) wait_flag(a.applied_flag, a.need_applied);
  __syncthreads();
  const long long nvec = a.n / 4;   // slices and tails are multiples of 8 elements
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = v * 4;      // index inside this server's slice
    const long long gi = a.lo + i;  // global index
    uint4 val;
    if (gi < a.total) {
      val = *reinterpret_cast<const uint4*>(a.grads + gi);
    } else {
      const long long r = gi - a.total;
      const float4 before = *reinterpret_cast<const float4*>(a.running_pulled + r);
      const float4 after = *reinterpret_cast<const float4*>(a.running + r);
      val.x = __float_as_uint(before.x - after.x), val.y = __float_as_uint(before.y - after.y);
      val.z = __float_as_uint(before.z - after.z), val.w = __float_as_uint(before.w - after.w);
    }
    st_relaxed_sys_v4(a.slot + i, val);
  }
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t is_last;
  if (threadIdx.x == 0) is_last = (atomicAdd(a.block_counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    *a.block_counter = 0;
    st_release_sys(a.ready_flag, a.seq);
  }
}

__global__ void __launch_bounds__(512) ps_pull_model_kernel(const PsPullArgs a) {
  const long long nvec = a.n / 8;
  for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = v * 8;
    const long long gi = a.lo + i;
    if (gi < a.total)   // bf16 serving copy: 8 weights per 16-byte load
      *reinterpret_cast<uint4*>(a.weights + gi) = ld_relaxed_sys_v4(a.wbf16 + i);
    if (gi >= a.decay_end) {  // fp32 tail: BN scale/offset + biases, then the running statistics
      const uint4 lo4 = ld_relaxed_sys_v4(a.master + i);
      const uint4 hi4 = ld_relaxed_sys_v4(a.master + i + 4);
      if (gi < a.total) {
        *reinterpret_cast<uint4*>(a.aux32 + (gi - a.decay_end)) = lo4;
        *reinterpret_cast<uint4*>(a.aux32 + (gi - a.decay_end) + 4) = hi4;
      } else {
        const long long r = gi - a.total;
        *reinterpret_cast<uint4*>(a.running + r) = lo4;
        *reinterpret_cast<uint4*>(a.running + r + 4) = hi4;
        *reinterpret_cast<uint4*>(a.running_pulled + r) = lo4;
        *reinterpret_cast<uint4*>(a.running_pulled + r + 4) = hi4;
      }
    }
  }
}

template <int OPT>
cudaError_t launch_ar(const AllreduceOptArgs& a, int grid, cudaStream_t s, int phase) {
  if (phase == 1)   // (the reduce-scatter half does not depend on the optimizer)
    allreduce_opt_kernel<kOptSgd, false, 1><<<grid, 512, 0, s>>>(a);
  else if (phase == 2)
    allreduce_opt_kernel<OPT, false, 2><<<grid, 512, 0, s>>>(a);
  else if (a.grads_mc != nullptr && a.weights_mc != nullptr && a.world > 1)
    allreduce_opt_kernel<OPT, true><<<grid, 512, 0, s>>>(a);
  else
    allreduce_opt_kernel<OPT, false><<<grid, 512, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t allreduce_opt(const AllreduceOptArgs& a, int opt, int grid, cudaStream_t s, int phase) {
  if (a.world < 1 || a.world > kMaxRanks) return cudaErrorInvalidValue;
  if ((a.begin & 7) != 0 || phase < 0 || phase > 2) return cudaErrorInvalidValue;
  switch (opt) {
    case kOptSgd: return launch_ar<kOptSgd>(a, grid, s, phase);
    case kOptMomentum: return launch_ar<kOptMomentum>(a, grid, s, phase);
    case kOptAdam: return launch_ar<kOptAdam>(a, grid, s, phase);
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t set_flag_timeout_ns(unsigned long long ns) {
  return cudaMemcpyToSymbol(g_flag_timeout_ns, &ns, sizeof(ns));
}
cudaError_t bcast_pull(const BcastArgs& a, int grid, cudaStream_t s) {
  bcast_pull_kernel<<<grid, 512, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t flag_barrier(const BcastArgs& a, cudaStream_t s) {
  flag_barrier_kernel<<<1, 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t ps_push_dense(float* w_ps, const float* g, long long n, const float* hyper,
                          cudaStream_t s) {
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  ps_push_dense_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(w_ps, g, n, hyper);
  return cudaGetLastError();
}
cudaError_t ps_push_sparse(float* w_ps, const float* g_rows, const int* idx, int nrows, int width,
                           const float* hyper, cudaStream_t s) {
  long long blocks = (static_cast<long long>(nrows) * width / ((width & 3) == 0 ? 4 : 1) + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  ps_push_sparse_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(w_ps, g_rows, idx, nrows,
                                                                     width, hyper);
  return cudaGetLastError();
}
cudaError_t ps_apply(const PsApplyArgs& a, int opt, cudaStream_t s) {
  if ((a.n & 7) != 0 || (a.lo & 7) != 0) return cudaErrorInvalidValue;
  const int grid = 148 * 2;
  switch (opt) {
    case kOptSgd: ps_apply_kernel<kOptSgd><<<grid, 512, 0, s>>>(a); break;
    case kOptMomentum: ps_apply_kernel<kOptMomentum><<<grid, 512, 0, s>>>(a); break;
    case kOptAdam: ps_apply_kernel<kOptAdam><<<grid, 512, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t ps_push_slot(const PsPushArgs& a, int grid, cudaStream_t s) {
  if ((a.n & 7) != 0 || (a.lo & 7) != 0 || (a.total & 7) != 0) return cudaErrorInvalidValue;
  ps_push_slot_kernel<<<grid, 512, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t ps_pull_model(const PsPullArgs& a, int grid, cudaStream_t s) {
  if ((a.n & 7) != 0 || (a.lo & 7) != 0 || (a.total & 7) != 0 || (a.decay_end & 7) != 0)
    return cudaErrorInvalidValue;
  ps_pull_model_kernel<<<grid, 512, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t ps_pull(const float* w_ps, float* w_local, void* w_bf16, long long n, cudaStream_t s) {
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  ps_pull_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(
      w_ps, w_local, static_cast<__nv_bfloat16*>(w_bf16), n);
  return cudaGetLastError();
}

}  // namespace tfos
