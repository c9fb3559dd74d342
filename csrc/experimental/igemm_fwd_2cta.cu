// DRAFT - compile-checked only (tools/compile_experimental.sh), never launched, not part of the
// shipped extension.  Round-2 starting point for DESIGN.md section 7 item 2.
//
// Why: profiles/ncu_l3c2_r1.txt shows that with cta_group::1 SS-mode MMAs the operand fetch from
// shared memory and the math add up (pipe active = math / (math + fetch) = 57 % for M128 x N256),
// and that the 3x3 layers are bound by re-reading B through the L2 -> SM fabric.  A CTA *pair*
// (cta_group::2, M = 256 across two SMs) lets each CTA fetch only HALF of B per MMA (4 KB A + 4 KB
// B for 128 clk of math per SM instead of 4 + 8) and halves the B traffic per output tile.
//
// Shape of the kernel (fprop-like problems, K-major B, N tile = 256, plain bf16 store epilogue):
//   cluster (2,1,1); CTA rank r of a pair owns pixel box 2*pair_tile + r  (its 128 rows of D)
//   smem stage per CTA : A box 128 x 64 (16 KB)  +  B rows [nt*256 + r*128, +128) x 64 (16 KB)
//   full[s]   : lives in the LEADER (rank 0); the TMA loads of BOTH CTAs complete_tx on it
//               (cp.async.bulk.tensor...cta_group::2 with the peer bit of the barrier address
//               cleared); the leader's producer arms it with the bytes of both CTAs
//   empty[s]  : one per CTA; the leader's tcgen05.commit.cta_group::2 multicasts to both
//   tfull[a]  : one per CTA (each epilogue reads its own TMEM half); multicast commit
//   tempty[a] : in the leader; the epilogue warps of both CTAs arrive on it (mapa for rank 1)
//   MMA       : leader's warp 1, one lane: tcgen05.mma.cta_group::2.kind::f16, idesc M = 256
//   TMEM      : tcgen05.alloc.cta_group::2 by warp 1 of both CTAs (same address in both)
// Open points to settle on hardware: the exact expect_tx accounting when both CTAs' TMA loads
// target the leader's barrier, and whether the 2-CTA descriptors want the leader's or the local
// shared-memory window for B (CUTLASS uses identical offsets in both CTAs - mirrored here).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../igemm.h"
#include "../ptx.cuh"

namespace tfos {
namespace experimental {

constexpr int kThreads2 = 320;           // producer warp, MMA warp, 8 epilogue warps
constexpr int kStages2 = 6;
constexpr int kA2 = 128 * 128;           // 128 pixels x 64 channels x 2 B
constexpr int kB2 = 128 * 128;           // half of the 256-row B tile
constexpr int kStage2 = kA2 + kB2;
constexpr int kSmem2 = kStages2 * kStage2 + 512;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address -> the leader's copy

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// TMA load whose completion is counted on the LEADER CTA's barrier (same smem offset, peer bit 0)
__device__ __forceinline__ void tma_load_4d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* m, uint64_t* bar,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive (count 1) on the same-offset barrier of every CTA in `mask` once the MMAs issued so far
// by this thread have completed
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

// tiles: pairs of 128-pixel boxes (tile_pairs of them) x n_tiles of 256 columns; 1x1 / dense
// problems only in this draft (one tap; a.k_chunks K blocks), tensor maps as in igemm.cu.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
igemm_fwd_2cta_kernel(const __grid_constant__ CUtensorMap tmA,
                      const __grid_constant__ CUtensorMap tmB, const FwdArgs a,
                      const int tile_pairs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages2 * kStage2);
  uint64_t* full = bars;                       // used in the leader only
  uint64_t* empty = bars + kStages2;           // one set per CTA
  uint64_t* tfull = bars + 2 * kStages2;       // [2] per CTA
  uint64_t* tempty = tfull + 2;                // [2] used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages2; ++i) {
      mbar_init(&full[i], 1);        // the leader's producer arms it (expect_tx) once per use
      mbar_init(&empty[i], 1);       // one multicast commit per use
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);     // 8 epilogue warps of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                // barriers + TMEM of both CTAs exist from here on
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = tile_pairs * a.n_tiles;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < total; t += npairs) {
        const int nt = t % a.n_tiles, mp = t / a.n_tiles;
        const int m0 = (2 * mp + static_cast<int>(rank)) * 128;      // this CTA's pixel box
        for (int kc = 0; kc < a.k_chunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStage2;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStage2);  // both CTAs' bytes
          tma_load_4d_2cta(sA, &tmA, &full[stage], kc * 64, m0, 0, 0);
          tma_load_2d_2cta(sA + kA2, &tmB, &full[stage], kc * 64,
                           nt * 256 + static_cast<int>(rank) * 128);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {    // the leader issues the MMAs of the pair
      const uint32_t idesc = umma_idesc_bf16(256, 256, false, false);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = pair; t < total; t += npairs) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kc = 0; kc < a.k_chunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * kStage2);
          const uint64_t ad0 = umma_desc_sw128(a_base, 16, 1024);
          const uint64_t bd0 = umma_desc_sw128(a_base + kA2, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2cta(d, ad0 + k * 2, bd0 + k * 2, idesc, (kc | k) != 0 ? 1u : 0u);
          umma_commit_2cta_mc(&empty[stage], 0x3);     // both producers may refill the stage
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2cta_mc(&tfull[acc], 0x3);          // both epilogues may read the tile
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // epilogue: this CTA's 128 rows x 256 columns (plain bf16 store; the production epilogue
    // of igemm.cu - staging, statistics, accumulate - slots in unchanged)
    const int q = warp & 3, half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
    for (int t = pair; t < total; t += npairs) {
      const int nt = t % a.n_tiles, mp = t / a.n_tiles;
      const long long row = static_cast<long long>(2 * mp + static_cast<int>(rank)) * 128 + q * 32 + lane;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = half * 4; c < half * 4 + 4; ++c) {    // 8 chunks of 32 columns, 4 per warp
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + c * 32, v);
        tmem_ld_wait();
        if (row < a.lim_w) {
          uint4* o = reinterpret_cast<uint4*>(out + row * a.ldo + nt * 256 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            o[j] = make_uint4(
                pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[acc], 0);   // the leader owns tempty
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 512);
}

cudaError_t launch_fwd_2cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const FwdArgs& a,
                            int tile_pairs, int num_sms, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(igemm_fwd_2cta_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2);
  if (e != cudaSuccess) return e;
  int pairs = num_sms / 2;
  if (pairs > tile_pairs * a.n_tiles) pairs = tile_pairs * a.n_tiles;
  igemm_fwd_2cta_kernel<<<2 * pairs, kThreads2, kSmem2, s>>>(tmA, tmB, a, tile_pairs);
  return cudaGetLastError();
}

}  // namespace experimental
}  // namespace tfos
