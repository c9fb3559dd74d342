// tcgen05 / TMEM / TMA implicit-GEMM kernels for sm_100a.
//
// Replaces what the reference reaches through TensorFlow -> cuDNN/cuBLAS for
// its example models (conv fprop/dgrad/wgrad, dense
// layers; reference call sites examples/mnist/keras/mnist_spark.py:13-20,
// examples/resnet/resnet_cifar_dist.py:208, examples/segmentation/
// segmentation_spark.py:67-119).
//
// Structure (both kernels): 192 threads = 6 warps, one CTA per SM, persistent
// over a static round-robin tile list.
//   warp 0   : TMA producer  (one lane issues cp.async.bulk.tensor into a smem ring)
//   warp 1   : TMEM allocator + MMA issuer (one lane issues tcgen05.mma; the
//              accumulator lives in TMEM, double-buffered across tiles)
//   warps 2-5: epilogue (tcgen05.ld -> registers -> bias/ReLU/stats -> global)
// Pipelines: smem full/empty mbarriers between TMA and MMA; TMEM full/empty
// mbarriers between MMA and epilogue, so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include <stdio.h>
#include <string.h>

#include <new>

#include "igemm.h"
#include "ptx.cuh"

namespace tfos {

namespace {

constexpr int kThreads = 192;
constexpr int kBlockM = 128;
constexpr int kABytes = kBlockM * 128;  // 128 rows x 64 bf16

template <int BN>
struct FwdCfg {
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 64) ? 8 : (BN <= 128 ? 6 : 4);
  static constexpr int kTmemCols = 2 * BN;  // double-buffered fp32 accumulator
  static constexpr int kStatBytes = 4 * BN * 2 * 4;
  static constexpr int kSmem = kStages * kStage + kStatBytes + 256 + 1024;
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
igemm_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const FwdArgs a, const int total_tiles) {
  using Cfg = FwdCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  float* stat_smem = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStage + Cfg::kStatBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_iters = a.num_taps * a.k_chunks;
  const int rows = a.box_w * a.box_h * a.box_n;
  const uint32_t stage_tx = static_cast<uint32_t>(rows) * 128u + Cfg::kBBytes;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_tiles;
        int mt = tile / a.n_tiles;
        const int tw = mt % a.tiles_w;
        mt /= a.tiles_w;
        const int th = mt % a.tiles_h;
        const int tn = mt / a.tiles_h;
        const int cw = tw * a.box_w * a.mul_w, ch = th * a.box_h * a.mul_h, cn = tn * a.box_n;
        for (int it = 0; it < k_iters; ++it) {
          const int t = it / a.k_chunks, kc = it - t * a.k_chunks;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStage;
          uint8_t* sB = sA + kABytes;
          mbar_expect_tx(&full[stage], stage_tx);
          tma_load_4d(sA, &tmA, &full[stage], kc * 64 + a.tap_dc[t], cw + a.tap_dw[t],
                      ch + a.tap_dh[t], cn);
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(sB + j * 8192, &tmB, &full[stage], nt * BN + j * 64 + a.tap_bn[t],
                          a.tap_bk[t] + kc * 64);
          } else {
            tma_load_2d(sB, &tmB, &full[stage], a.tap_bk[t] + kc * 64, nt * BN);
          }
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, false, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::kStage);
          const uint32_t b_base = a_base + kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = umma_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : umma_desc_sw128(b_base + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int ew = warp - 2;
    const int et = threadIdx.x - 64;  // 0..127
    const bool do_stats = a.col_sum != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % a.n_tiles;
      int mt = tile / a.n_tiles;
      const int tw = mt % a.tiles_w;
      mt /= a.tiles_w;
      const int th = mt % a.tiles_h;
      const int tn = mt / a.tiles_h;
      const int row = q * 32 + lane;
      const int iw = row % a.box_w;
      const int ih = (row / a.box_w) % a.box_h;
      const int in_ = row / (a.box_w * a.box_h);
      const int w = tw * a.box_w + iw, h = th * a.box_h + ih, n = tn * a.box_n + in_;
      const bool valid = row < rows && w < a.lim_w && h < a.lim_h && n < a.lim_n;
      const long long pix =
          (static_cast<long long>(n) * a.OH + (h * a.osh + a.ooh)) * a.OW + (w * a.osw + a.oow);
      const long long off = pix * a.ldo + static_cast<long long>(nt) * BN;

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = nt * BN + c * 32;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (a.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < a.n_valid) f[j] += __ldg(a.bias + col0 + j);
        }
        if (a.out_fp32) {
          float* o = reinterpret_cast<float*>(a.out) + off + c * 32;
          if (valid) {
            if (a.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < a.n_valid) f[j] += o[j];
            }
            if (a.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (col0 + 32 <= a.n_valid && (a.ldo & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < a.n_valid) o[j] = f[j];
            }
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(a.out) + off + c * 32;
          const bool vec = col0 + 32 <= a.n_valid && (a.ldo & 7) == 0;
          if (valid && a.accumulate) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 p = *reinterpret_cast<const uint4*>(o + j);
                const float2 p0 = unpack_bf16x2(p.x), p1 = unpack_bf16x2(p.y),
                             p2 = unpack_bf16x2(p.z), p3 = unpack_bf16x2(p.w);
                f[j] += p0.x, f[j + 1] += p0.y, f[j + 2] += p1.x, f[j + 3] += p1.y;
                f[j + 4] += p2.x, f[j + 5] += p2.y, f[j + 6] += p3.x, f[j + 7] += p3.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < a.n_valid) f[j] += __bfloat162float(o[j]);
            }
          }
          if (a.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          // Round once; statistics are taken of the values actually stored.
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __bfloat162float(__float2bfloat16_rn(f[j]));
          if (valid) {
            if (vec) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 p;
                p.x = pack_bf16x2(f[j], f[j + 1]);
                p.y = pack_bf16x2(f[j + 2], f[j + 3]);
                p.z = pack_bf16x2(f[j + 4], f[j + 5]);
                p.w = pack_bf16x2(f[j + 6], f[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = p;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < a.n_valid) o[j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
        if (do_stats) {
          // Column sums over the 32 rows of this warp by a transpose-reduce
          // butterfly: after 5 steps lane j holds the totals of column j.
          float s[32], ss[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = valid ? f[j] : 0.f;
            s[j] = x;
            ss[j] = x * x;
          }
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const bool upper = (lane & step) != 0;
#pragma unroll
            for (int j = 0; j < step; ++j) {
              const float send_s = upper ? s[j] : s[j + step];
              const float send_q = upper ? ss[j] : ss[j + step];
              const float recv_s = __shfl_xor_sync(0xffffffff, send_s, step);
              const float recv_q = __shfl_xor_sync(0xffffffff, send_q, step);
              s[j] = (upper ? s[j + step] : s[j]) + recv_s;
              ss[j] = (upper ? ss[j + step] : ss[j]) + recv_q;
            }
          }
          stat_smem[(ew * BN + c * 32 + lane) * 2 + 0] = s[0];
          stat_smem[(ew * BN + c * 32 + lane) * 2 + 1] = ss[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (do_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int c = et; c < BN; c += 128) {
          const int col = nt * BN + c;
          if (col < a.n_valid) {
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int wq = 0; wq < 4; ++wq) {
              s += stat_smem[(wq * BN + c) * 2 + 0];
              ss += stat_smem[(wq * BN + c) * 2 + 1];
            }
            atomicAdd(a.col_sum + col, s);
            atomicAdd(a.col_sumsq + col, ss);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------- wgrad
// dW[m, tap, n] += sum_pixels dY[pixel, m] * X[pixel + tap, n].  Both operands
// are MN-major in shared memory (the contiguous global dimension is channels,
// the reduction runs over pixel rows), one TMA box of <=128 pixels per stage.
template <int BN>
struct WgCfg {
  static constexpr int kAStage = 2 * 128 * 128;          // two 64-channel boxes, <=128 pixels
  static constexpr int kBStage = (BN / 64) * 128 * 128;
  static constexpr int kStage = kAStage + kBStage;
  static constexpr int kStages = (BN <= 64) ? 4 : (BN <= 128 ? 3 : 2);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmem = kStages * kStage + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
igemm_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const WgradArgs a, const int total_work) {
  using Cfg = WgCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_boxes = a.tiles_w * a.tiles_h * a.tiles_n;
  const int boxes_per_split = (total_boxes + a.k_splits - 1) / a.k_splits;
  const uint32_t box_bytes = static_cast<uint32_t>(a.box_rows) * 128u;
  const uint32_t stage_tx = box_bytes * (2 + BN / 64);

  // work item -> (tap, m tile, n tile, k split); splits vary fastest so that
  // concurrently running CTAs stream disjoint pixels of the same tile.
  auto decode = [&](int work, int& t, int& mt, int& nt, int& b0, int& b1) {
    const int ks = work % a.k_splits;
    int r = work / a.k_splits;
    nt = r % a.n_tiles;
    r /= a.n_tiles;
    mt = r % a.m_tiles;
    t = r / a.m_tiles;
    b0 = ks * boxes_per_split;
    b1 = min(b0 + boxes_per_split, total_boxes);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int t, mt, nt, b0, b1;
        decode(work, t, mt, nt, b0, b1);
        for (int b = b0; b < b1; ++b) {
          const int tw = b % a.tiles_w;
          const int th = (b / a.tiles_w) % a.tiles_h;
          const int tn = b / (a.tiles_w * a.tiles_h);
          const int pw = tw * a.box_w, ph = th * a.box_h, pn = tn * a.box_n;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStage;
          uint8_t* sB = sA + Cfg::kAStage;
          mbar_expect_tx(&full[stage], stage_tx);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_4d(sA + j * box_bytes, &tmA, &full[stage], mt * 128 + j * 64, pw, ph, pn);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_4d(sB + j * box_bytes, &tmB, &full[stage], nt * BN + j * 64 + a.tap_dc[t],
                        pw * a.mul_w + a.tap_dw[t], ph * a.mul_h + a.tap_dh[t], pn);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int mmas = a.box_rows / 16;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int t, mt, nt, b0, b1;
        decode(work, t, mt, nt, b0, b1);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int b = b0; b < b1; ++b) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::kStage);
          const uint32_t b_base = a_base + Cfg::kAStage;
          for (int k = 0; k < mmas; ++k) {
            const uint64_t adesc = umma_desc_sw128(a_base + k * 2048, box_bytes, 1024);
            const uint64_t bdesc = umma_desc_sw128(b_base + k * 2048, box_bytes, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (b > b0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      int t, mt, nt, b0, b1;
      decode(work, t, mt, nt, b0, b1);
      const int m = mt * 128 + q * 32 + lane;
      const bool valid = m < a.m_valid && b1 > b0;
      float* o = a.dw + static_cast<long long>(m) * a.ldw + a.tap_out[t] + nt * BN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col0 = nt * BN + c * 32;
        if (valid) {
          if (col0 + 32 <= a.n_valid && (a.ldw & 3) == 0 && (a.tap_out[t] & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_f32x4(o + c * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < a.n_valid) atomicAdd(o + c * 32 + j, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// -------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool encode(const TmapDesc& d, CUtensorMap* out, char* err, int errlen) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return false;
  }
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], es[4];
  for (int i = 0; i < d.rank; ++i) {
    dims[i] = d.dims[i];
    box[i] = d.box[i];
    es[i] = d.elem_strides[i];
  }
  for (int i = 0; i + 1 < d.rank; ++i) strides[i] = d.strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, d.rank, d.base, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen,
             "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] "
             "strides=[%llu,%llu,%llu] box=[%u,%u,%u,%u] es=[%u,%u,%u,%u] base=%p",
             static_cast<int>(r), d.rank, (unsigned long long)d.dims[0],
             (unsigned long long)d.dims[1], (unsigned long long)d.dims[2],
             (unsigned long long)d.dims[3], (unsigned long long)d.strides_bytes[0],
             (unsigned long long)d.strides_bytes[1], (unsigned long long)d.strides_bytes[2],
             d.box[0], d.box[1], d.box[2], d.box[3], d.elem_strides[0], d.elem_strides[1],
             d.elem_strides[2], d.elem_strides[3], d.base);
    return false;
  }
  return true;
}

template <typename K>
cudaError_t set_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <int BN, bool B_MN>
cudaError_t launch_fwd(const IGemmPlan* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = set_smem(igemm_fwd_kernel<BN, B_MN>, FwdCfg<BN>::kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  igemm_fwd_kernel<BN, B_MN><<<p->grid, kThreads, FwdCfg<BN>::kSmem, s>>>(p->tmA, p->tmB, p->fa,
                                                                           p->total_work);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_wgrad(const IGemmPlan* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = set_smem(igemm_wgrad_kernel<BN>, WgCfg<BN>::kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  igemm_wgrad_kernel<BN><<<p->grid, kThreads, WgCfg<BN>::kSmem, s>>>(p->tmA, p->tmB, p->wa,
                                                                    p->total_work);
  return cudaGetLastError();
}

}  // namespace

IGemmPlan* igemm_plan_fwd(const TmapDesc& a, const TmapDesc& b, const FwdArgs& args, int bn,
                          int b_mn, int num_sms, char* err, int errlen) {
  if (bn != 64 && bn != 128 && bn != 256) {
    snprintf(err, errlen, "bn must be 64/128/256");
    return nullptr;
  }
  const int rows = args.box_w * args.box_h * args.box_n;
  if (rows < 1 || rows > 128 || args.num_taps < 1 || args.num_taps > kMaxTaps) {
    snprintf(err, errlen, "bad box rows %d or taps %d", rows, args.num_taps);
    return nullptr;
  }
  IGemmPlan* p = new (std::nothrow) IGemmPlan();
  if (!p) return nullptr;
  memset(p, 0, sizeof(*p));
  if (!encode(a, &p->tmA, err, errlen) || !encode(b, &p->tmB, err, errlen)) {
    delete p;
    return nullptr;
  }
  p->fa = args;
  p->kind = 0;
  p->bn = bn;
  p->b_mn = b_mn;
  p->total_work = args.n_tiles * args.tiles_w * args.tiles_h * args.tiles_n;
  p->grid = p->total_work < num_sms ? p->total_work : num_sms;
  return p;
}

IGemmPlan* igemm_plan_wgrad(const TmapDesc& a, const TmapDesc& b, const WgradArgs& args, int bn,
                            int num_sms, char* err, int errlen) {
  if (bn != 64 && bn != 128 && bn != 256) {
    snprintf(err, errlen, "bn must be 64/128/256");
    return nullptr;
  }
  if (args.box_rows % 16 != 0 || args.box_rows > 128 || args.box_rows < 16 ||
      args.box_rows != args.box_w * args.box_h * args.box_n) {
    snprintf(err, errlen, "wgrad box_rows %d must be a multiple of 16, <= 128", args.box_rows);
    return nullptr;
  }
  IGemmPlan* p = new (std::nothrow) IGemmPlan();
  if (!p) return nullptr;
  memset(p, 0, sizeof(*p));
  if (!encode(a, &p->tmA, err, errlen) || !encode(b, &p->tmB, err, errlen)) {
    delete p;
    return nullptr;
  }
  p->wa = args;
  p->kind = 1;
  p->bn = bn;
  p->total_work = args.num_taps * args.m_tiles * args.n_tiles * args.k_splits;
  p->grid = p->total_work < num_sms ? p->total_work : num_sms;
  return p;
}

cudaError_t igemm_run(const IGemmPlan* p, cudaStream_t s) {
  if (p->kind == 0) {
    if (p->b_mn) {
      switch (p->bn) {
        case 64: return launch_fwd<64, true>(p, s);
        case 128: return launch_fwd<128, true>(p, s);
        default: return launch_fwd<256, true>(p, s);
      }
    }
    switch (p->bn) {
      case 64: return launch_fwd<64, false>(p, s);
      case 128: return launch_fwd<128, false>(p, s);
      default: return launch_fwd<256, false>(p, s);
    }
  }
  switch (p->bn) {
    case 64: return launch_wgrad<64>(p, s);
    case 128: return launch_wgrad<128>(p, s);
    default: return launch_wgrad<256>(p, s);
  }
}

void igemm_plan_free(IGemmPlan* p) { delete p; }

}  // namespace tfos
