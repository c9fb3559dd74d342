// Weight-gradient tcgen05 kernel (split out of igemm.cu): dW[m, tap, n] += sum over pixels.
// See igemm.cu for the structure shared by both kernels (TMA producer warp, single-thread
// MMA issuer with TMEM accumulators, epilogue warps).
#include <stdio.h>
#include <string.h>

#include "igemm.h"
#include "ptx.cuh"

namespace tfos {

namespace {

constexpr int kThreads = 192;
constexpr int kBlockM = 128;

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

// ------------------------------------------------------------- stem wgrad
// 7x7 / stride-2 stem on the window-row layout (ops/igemm.py): dW[m, r, 0..63] for all seven
// filter rows r from ONE pass over the pixels.  Per stage of 16 x 8 output pixels:
//   A = dY box  [128 pixels][64 channels]              16 KB  (MN-major, M = Cout = 64)
//   B = X  box  [21 input rows x 16 window columns][64] 42 KB  (MN-major, N = 64)
// Output row j (16 pixels = one K=16 slice = 2048 B) of filter row r multiplies box row 2j + r,
// so the seven filter rows differ only in the B descriptor's start address; they accumulate
// into seven 64-column TMEM accumulators (448 of 512 columns).  Compared with one pass per
// filter row (the generic kernel) the L2 -> SM traffic drops 3.9x, which is what bounded it.
// The MMA is M = 128 wide; rows 64..127 of every accumulator are scratch (their A "half" is
// whatever follows the dY box in shared memory) and are never read.
constexpr int kStemBW = 16, kStemBH = 8, kStemXRows = 2 * kStemBH + 5;
constexpr int kStemA = kStemBW * kStemBH * 128;      // 16384
constexpr int kStemB = kStemBW * kStemXRows * 128;   // 43008
constexpr int kStemStage = kStemA + kStemB;          // 59392 (multiple of 1024)
constexpr int kStemStages = 3;
constexpr int kStemSmem = kStemStages * kStemStage + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
igemm_wgrad_stem_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB, const WgradArgs a,
                        const int total_work) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStemStages * kStemStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStemStages;
  uint64_t* tfull = bars + 2 * kStemStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStemStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_boxes = a.tiles_w * a.tiles_h * a.tiles_n;
  const int per_split = (total_boxes + a.k_splits - 1) / a.k_splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int b0 = work * per_split, b1 = min(b0 + per_split, total_boxes);
        for (int b = b0; b < b1; ++b) {
          const int tw = b % a.tiles_w;
          const int th = (b / a.tiles_w) % a.tiles_h;
          const int tn = b / (a.tiles_w * a.tiles_h);
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStemStage;
          mbar_expect_tx(&full[stage], kStemStage);
          tma_load_4d(sA, &tmA, &full[stage], 0, tw * kStemBW, th * kStemBH, tn);
          tma_load_4d(sA + kStemA, &tmB, &full[stage], 0, tw * kStemBW + a.tap_dw[0],
                      th * kStemBH * 2 + a.tap_dh[0], tn);
          if (++stage == kStemStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, 64, true, true);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int b0 = work * per_split, b1 = min(b0 + per_split, total_boxes);
        mbar_wait(tempty, acc_phase ^ 1);
        tc_fence_after();
        for (int b = b0; b < b1; ++b) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * kStemStage);
          const uint32_t b_base = a_base + kStemA;
          for (int r = 0; r < 7; ++r) {
#pragma unroll
            for (int j = 0; j < kStemBH; ++j) {
              // A: second 64-row half = kStemA bytes further on (scratch rows, see above)
              const uint64_t adesc = umma_desc_sw128(a_base + j * 2048, kStemA, 1024);
              const uint64_t bdesc = umma_desc_sw128(b_base + (2 * j + r) * 2048, 2048, 1024);
              umma_bf16(tmem_base + r * 64, adesc, bdesc, idesc, (b > b0 || j > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == kStemStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull);
        acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const int b0 = work * per_split, b1 = min(b0 + per_split, total_boxes);
      const int m = q * 32 + lane;
      const bool valid = m < a.m_valid && m < 64 && b1 > b0;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      if (q < 2) {
#pragma unroll 1
        for (int c = 0; c < 14; ++c) {  // 7 filter rows x two 32-column chunks
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          if (valid) {
            float* o = a.dw + static_cast<long long>(m) * a.ldw + c * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_f32x4(o + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------- wide wgrad
// 256 x 256 output tile per CTA (two 128-row accumulators x 256 columns = all 512 TMEM columns):
// the generic kernel's 128 x 128 tile needs (128+128) channels x 2 B of operands per pixel for
// 16 K MACs and is bound by the L2 -> SM fabric (profiles/ncu_l3c2_r1.txt: 10.3 TB/s of L2 reads at
// 33 % tensor activity); this tile needs (256+256) x 2 B for 64 K MACs - half the bytes per MAC.
// Stages are 64 pixels (4 + 4 boxes of 8 KB) so that three of them fit; the accumulators are
// single-buffered (a work item's K loop is long, its epilogue short).
constexpr int kWideBox = 64 * 128;                  // 64 pixels x 64 channels x 2 B
constexpr int kWideStage = 8 * kWideBox;            // A: 2 m-tiles x 2 halves, B: 4 column blocks
constexpr int kWideStages = 3;
constexpr int kWideSmem = kWideStages * kWideStage + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1)
igemm_wgrad_wide_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB, const WgradArgs a,
                        const int total_work) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWideStages * kWideStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWideStages;
  uint64_t* tfull = bars + 2 * kWideStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kWideStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_boxes = a.tiles_w * a.tiles_h * a.tiles_n;
  const int boxes_per_split = (total_boxes + a.k_splits - 1) / a.k_splits;
  const uint32_t box_bytes = static_cast<uint32_t>(a.box_rows) * 128u;   // <= kWideBox
  const uint32_t stage_tx = box_bytes * 8u;
  const int m_pairs = (a.m_tiles + 1) / 2;

  auto decode = [&](int work, int& t, int& mp, int& nt, int& b0, int& b1) {
    const int ks = work % a.k_splits;
    int r = work / a.k_splits;
    nt = r % a.n_tiles;
    r /= a.n_tiles;
    mp = r % m_pairs;
    t = r / m_pairs;
    b0 = ks * boxes_per_split;
    b1 = min(b0 + boxes_per_split, total_boxes);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int t, mp, nt, b0, b1;
        decode(work, t, mp, nt, b0, b1);
        for (int b = b0; b < b1; ++b) {
          const int tw = b % a.tiles_w;
          const int th = (b / a.tiles_w) % a.tiles_h;
          const int tn = b / (a.tiles_w * a.tiles_h);
          const int pw = tw * a.box_w, ph = th * a.box_h, pn = tn * a.box_n;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kWideStage;
          uint8_t* sB = sA + 4 * kWideBox;
          mbar_expect_tx(&full[stage], stage_tx);
#pragma unroll
          for (int j = 0; j < 4; ++j)   // channels beyond Cout are zero-filled by TMA
            tma_load_4d(sA + j * box_bytes, &tmA, &full[stage], mp * 256 + j * 64, pw, ph, pn);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_4d(sB + j * box_bytes, &tmB, &full[stage], nt * 256 + j * 64 + a.tap_dc[t],
                        pw * a.mul_w + a.tap_dw[t], ph * a.mul_h + a.tap_dh[t], pn);
          if (++stage == kWideStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, 256, true, true);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const int mmas = a.box_rows / 16;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int t, mp, nt, b0, b1;
        decode(work, t, mp, nt, b0, b1);
        mbar_wait(tempty, acc_phase ^ 1);
        tc_fence_after();
        for (int b = b0; b < b1; ++b) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * kWideStage);
          const uint32_t b_base = a_base + 4 * kWideBox;
          for (int k = 0; k < mmas; ++k) {
            const uint64_t bdesc = umma_desc_sw128(b_base + k * 2048, box_bytes, 1024);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint64_t adesc =
                  umma_desc_sw128(a_base + mt * 2 * box_bytes + k * 2048, box_bytes, 1024);
              umma_bf16(tmem_base + mt * 256, adesc, bdesc, idesc, (b > b0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == kWideStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull);
        acc_phase ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      int t, mp, nt, b0, b1;
      decode(work, t, mp, nt, b0, b1);
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
      const bool vec_ok = (a.ldw & 3) == 0 && (a.tap_out[t] & 3) == 0;
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int m = mp * 256 + mt * 128 + q * 32 + lane;
        const bool valid = m < a.m_valid && b1 > b0;
        float* o = a.dw + static_cast<long long>(m) * a.ldw + a.tap_out[t] + nt * 256;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + mt * 256 + c * 32, v);
          tmem_ld_wait();
          const int col0 = nt * 256 + c * 32;
          if (valid) {
            if (col0 + 32 <= a.n_valid && vec_ok) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

cudaError_t launch_wgrad_wide(const IGemmPlan* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_wide_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kWideSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  igemm_wgrad_wide_kernel<<<p->grid, kThreads, kWideSmem, s>>>(p->tmA, p->tmB, p->wa,
                                                              p->total_work);
  return cudaGetLastError();
}

cudaError_t launch_wgrad_stem(const IGemmPlan* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_stem_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  igemm_wgrad_stem_kernel<<<p->grid, kThreads, kStemSmem, s>>>(p->tmA, p->tmB, p->wa,
                                                              p->total_work);
  return cudaGetLastError();
}

template <int BN>
cudaError_t launch_wgrad(const IGemmPlan* p, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(igemm_wgrad_kernel<BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WgCfg<BN>::kSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  igemm_wgrad_kernel<BN><<<p->grid, kThreads, WgCfg<BN>::kSmem, s>>>(p->tmA, p->tmB, p->wa,
                                                                    p->total_work);
  return cudaGetLastError();
}

}  // namespace

cudaError_t igemm_run_wgrad(const IGemmPlan* p, cudaStream_t s) {
  if (p->wa.stem) return launch_wgrad_stem(p, s);
  if (p->wa.wide) return launch_wgrad_wide(p, s);
  switch (p->bn) {
    case 64: return launch_wgrad<64>(p, s);
    case 128: return launch_wgrad<128>(p, s);
    default: return launch_wgrad<256>(p, s);
  }
}

}  // namespace tfos
