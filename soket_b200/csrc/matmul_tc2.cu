// matmul_tc2.cu -- the GEMM of matmul_tc.cu on CTA PAIRS: tcgen05.mma.cta_group::2.
//
// Two CTAs of a cluster (two SMs of one TPC) compute one 256 x 256 output tile.  Each
// CTA loads its own 128 rows of A and HALF of the B tile (128 of the 256 N-columns)
// into its own shared memory; the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256), which reads A from each CTA's smem and the two
// B halves from both, and accumulates rows 0-127 in the leader's TMEM, rows 128-255 in
// the peer's.  Per CTA the shared-memory traffic per flop halves for B compared with
// the single-CTA kernel, which is what bounds that kernel.
//
// Synchronisation (all mbarriers live at identical offsets in both CTAs):
//   full[s]   (leader's copy, count 1): the leader's producer arrives with expect_tx of
//             BOTH CTAs' bytes; both CTAs' TMA loads (.cta_group::2) complete_tx on the
//             leader's barrier (bytes-only handshake, no remote arrive by the peer)
//   empty[s]  (own copy, count 1): tcgen05.commit.cta_group::2 ... multicast 0b11
//   tfull[a]  (own copy, count 1): multicast commit -> each CTA's epilogue
//   tempty[a] (leader's copy, count 2 x epilogue warps): every epilogue warp of both
//             CTAs arrives (the peer's remotely)
//
// Tile scheduling is DYNAMIC: a scheduler warp in the leader CTA hands out tile indices
// from a global counter (the first tile of each pair is static) through a 2-deep ring that
// exists in both CTAs (sfull[i] per CTA; sempty[i] on the leader, count = every reader of
// both CTAs).  A pair whose SMs were busy when the kernel started -- an NCCL all-reduce of
// the previous layer's gradients running on the comm stream -- simply claims fewer tiles,
// where the static `tile += pairs` order made the whole GEMM wait for its slowest pair
// (measured: +3.5 ms of GEMM time per step at 8 GPUs).
#include <stdlib.h>

#include "common.cuh"
#include "matmul.cuh"
#include "matmul_tc.cuh"

namespace sk {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Remote arrive WITHOUT release semantics: TMEM reads are ordered by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync; a .release here would wait for the epilogue's own
// global stores of the previous tile.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"((uint16_t)3)
      : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
// wait with cluster-scope acquire: the value published next to the barrier came from the other CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

constexpr int BM2 = 256;          // rows per CTA pair
constexpr int BMH = 128;          // rows per CTA
constexpr int BN2 = 256;          // columns per pair tile; each CTA stages BN2/2 of B
constexpr int kTc2Threads = 64 + 128 * (BN2 / 128) + 32;   // TMA, MMA, 8 epilogue warps, scheduler
constexpr int SD = 2;             // tile-ring depth: tiles a pair may claim ahead

template <int KIND, int STAGES, int CHUNK_KB, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTc2Threads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_alo,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_blo,
                const TcParams p) {
  constexpr bool BF16 = KIND == KIND_BF16 || KIND == KIND_F16X3;   // 16-bit operands (kind::f16)
  constexpr bool X3 = KIND == KIND_TF32X3 || KIND == KIND_F16X3;   // hi/lo operand pairs, 3 MMAs per K step
  constexpr int ES = BF16 ? 2 : 4;
  constexpr int BK = 128 / ES;
  constexpr int UK = 32 / ES;
  constexpr int SLAB = 128 / ES;
  constexpr int BNH = BN2 / 2;                    // B columns staged by this CTA
  constexpr uint32_t A_BYTES = BMH * 128;
  constexpr uint32_t B_BYTES = BNH * 128;
  constexpr uint32_t STAGE_BYTES = (X3 ? 2 : 1) * (A_BYTES + B_BYTES);   // per CTA
  constexpr uint32_t A_LO_OFF = A_BYTES;
  constexpr uint32_t B_OFF = (X3 ? 2 : 1) * A_BYTES;
  constexpr uint32_t B_LO_OFF = B_OFF + B_BYTES;
  constexpr uint32_t A_LBO = A_MN ? BK * 128 : 0, A_SBO = (A_MN && !BF16) ? 512 : 1024;
  constexpr uint32_t B_LBO = B_MN ? BK * 128 : 0, B_SBO = (B_MN && !BF16) ? 512 : 1024;
  constexpr uint32_t A_LT = (A_MN && !BF16) ? 1 : 2, B_LT = (B_MN && !BF16) ? 1 : 2;
  constexpr uint32_t A_KSTEP = A_MN ? UK * 128 : 32;
  constexpr uint32_t B_KSTEP = B_MN ? UK * 128 : 32;
  constexpr uint32_t IDESC = make_idesc(KIND == KIND_F16X3 ? 0 : (BF16 ? 1 : 2), A_MN, B_MN, BM2, BN2);
  constexpr int TMEM_COLS = 2 * BN2;              // 512: two accumulator buffers
  constexpr int EPI_WARPS = 4 * (BN2 / 128);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t *full_bar = bars;
  uint64_t *empty_bar = bars + STAGES;
  uint64_t *tfull_bar = bars + 2 * STAGES;
  uint64_t *tempty_bar = bars + 2 * STAGES + 2;
  uint64_t *sfull_bar = bars + 2 * STAGES + 4;      // [SD] tile id published (own copy)
  uint64_t *sempty_bar = bars + 2 * STAGES + 4 + SD;  // [SD] tile id read by everyone (leader's copy)
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 4 + 2 * SD);
  volatile int *tile_ring = (volatile int *)(tmem_slot + 2);   // [SD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = p.tiles_m * p.tiles_n;      // tiles of 256 x 256
  const int num_kb = (p.K + BK - 1) / BK;
  const bool dyn = p.sched != nullptr;
  // every reader of a ring slot: TMA thread + epilogue warps of both CTAs + the MMA thread
  constexpr int kRingReaders = 2 * (1 + EPI_WARPS) + 1;

  // The it-th tile of this pair, or -1.  Called by one lane (TMA / MMA threads) or by a whole
  // warp (`warp_wide`, epilogue); `elected` (one lane per caller) releases the ring slot once
  // every calling lane has the value.
  int sched_it = 0;
  auto next_tile = [&](bool elected, bool warp_wide) -> int {
    int t;
    if (!dyn) {
      t = pair + sched_it * num_pairs;
      if (t >= num_tiles) t = -1;
    } else {
      const int sl = sched_it % SD;
      mbar_wait_cluster(smem_u32(&sfull_bar[sl]), (uint32_t)((sched_it / SD) & 1));
      t = tile_ring[sl];
      if (warp_wide) __syncwarp();
      if (elected && t >= -1) {   // the (always true) test makes the arrive depend on the loaded value
        if (rank == 0) mbar_arrive(smem_u32(&sempty_bar[sl]));
        else mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(&sempty_bar[sl]), 0));
      }
    }
    ++sched_it;
    return t;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (X3) { tma_prefetch_desc(&map_alo); tma_prefetch_desc(&map_blo); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);         // the leader's expect_tx covers both CTAs' bytes
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 2 * EPI_WARPS);
    }
    for (int i = 0; i < SD; ++i) {
      mbar_init(smem_u32(&sfull_bar[i]), 1);
      mbar_init(smem_u32(&sempty_bar[i]), kRingReaders);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                               // barriers + TMEM visible in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = next_tile(true, false); tile >= 0; tile = next_tile(true, false)) {
        int tm, tn;
        tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
        const int m0 = tm * BM2 + (int)rank * BMH;       // this CTA's A rows
        const int n0 = tn * BN2 + (int)rank * BNH;       // this CTA's half of B
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb_leader = map_to_cta(smem_u32(&full_bar[stage]), 0);
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          // Bytes-only handshake: the peer does NOT arrive on the leader's barrier.  A remote
          // mbarrier.arrive.release.cluster from this thread would first wait for its own
          // outstanding TMA writes, serialising the ring to one TMA latency per stage
          // (measured: 337 -> 675 TFLOP/s for single-pass TF32 when removed).
          if (rank == 0) mbar_expect_tx(smem_u32(&full_bar[stage]), 2 * STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BMH / SLAB; ++j) {
              tma_load_2d_2sm(sbase + j * (BK * 128), &map_a, fb_leader, m0 + j * SLAB, k0);
              if (X3) tma_load_2d_2sm(sbase + A_LO_OFF + j * (BK * 128), &map_alo, fb_leader, m0 + j * SLAB, k0);
            }
          } else {
            tma_load_2d_2sm(sbase, &map_a, fb_leader, k0, m0);
            if (X3) tma_load_2d_2sm(sbase + A_LO_OFF, &map_alo, fb_leader, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BNH / SLAB; ++j) {
              tma_load_2d_2sm(sbase + B_OFF + j * (BK * 128), &map_b, fb_leader, n0 + j * SLAB, k0);
              if (X3) tma_load_2d_2sm(sbase + B_LO_OFF + j * (BK * 128), &map_blo, fb_leader, n0 + j * SLAB, k0);
            }
          } else {
            tma_load_2d_2sm(sbase + B_OFF, &map_b, fb_leader, k0, n0);
            if (X3) tma_load_2d_2sm(sbase + B_LO_OFF, &map_blo, fb_leader, k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = next_tile(true, false); tile >= 0; tile = next_tile(true, false)) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB) {
          const int kb1 = kb0 + CHUNK_KB < num_kb ? kb0 + CHUNK_KB : num_kb;
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);   // both CTAs drained this buffer
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN2);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);          // both CTAs' bytes have landed
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da = make_smem_desc(sbase + k * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t db = make_smem_desc(sbase + B_OFF + k * B_KSTEP, B_LBO, B_SBO, B_LT);
              tc_mma_2sm<BF16>(d_tmem, da, db, IDESC, ((kb - kb0) | k) != 0);
              if (X3) {
                const uint64_t dalo = make_smem_desc(sbase + A_LO_OFF + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dblo = make_smem_desc(sbase + B_LO_OFF + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                tc_mma_2sm<BF16>(d_tmem, da, dblo, IDESC, 1);
                tc_mma_2sm<BF16>(d_tmem, dalo, db, IDESC, 1);
              }
            }
            tc_commit_2sm(smem_u32(&empty_bar[stage]));            // frees the slot in BOTH CTAs
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit_2sm(smem_u32(&tfull_bar[acc]));                // wakes both epilogues
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp == 2 + EPI_WARPS) {
    // ===================== tile scheduler (leader CTA only) =====================
    if (dyn && rank == 0 && lane == 0) {
      for (int it = 0;; ++it) {
        const int sl = it % SD;
        mbar_wait_cluster(smem_u32(&sempty_bar[sl]), (uint32_t)(((it / SD) & 1) ^ 1));   // all 19 readers done
        int t = it == 0 ? pair : num_pairs + atomicAdd(p.sched, 1);
        if (t >= num_tiles) t = -1;
        tile_ring[sl] = t;
        st_cluster_u32(map_to_cta(smem_u32((const void *)&tile_ring[sl]), 1), (uint32_t)t);
        mbar_arrive(smem_u32(&sfull_bar[sl]));
        mbar_arrive_cluster_release(map_to_cta(smem_u32(&sfull_bar[sl]), 1));
        if (t < 0) break;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int q = warp & 3;
    const int cbase = ((warp - 2) >> 2) * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.ldc % 4 == 0) && ((((uintptr_t)p.c) & 15) == 0);
    for (int tile = next_tile(lane == 0, true); tile >= 0; tile = next_tile(lane == 0, true)) {
      int tm, tn;
      tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, tm, tn);
      const int m0 = tm * BM2 + (int)rank * BMH;
      const int n0 = tn * BN2 + cbase;
      float accum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) accum[j] = 0.f;
      const bool single = num_kb <= CHUNK_KB;
      for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
        if (n0 < p.N) {
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2 + cbase + c0), r);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              accum[c0 + j] = single ? __uint_as_float(r[j]) : __fadd_rn(accum[c0 + j], __uint_as_float(r[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
          else mbar_arrive_cluster_relaxed(map_to_cta(smem_u32(&tempty_bar[acc]), 0));
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int row = m0 + q * 32 + lane;
      if (row < p.M && n0 < p.N) {
        float *crow = p.c + (int64_t)row * p.ldc;
        const float rinv = (KIND != KIND_F16X3) ? 1.f
                           : p.row_inv ? __ldg(p.row_inv + row) : (p.a_inv1 ? __ldg(p.a_inv1) : 1.f);
        const float cinv1 = (KIND == KIND_F16X3 && p.b_inv1) ? __ldg(p.b_inv1) : 1.f;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          const int col = n0 + c0;
          if (col < p.N) {
            const bool full = col + 32 <= p.N;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = accum[c0 + j];
              if (KIND == KIND_F16X3) {
                // undo the operand scales: exact powers of two, smaller factor first so that the
                // intermediate cannot overflow when the result itself is representable
                const float cinv = p.col_inv ? ((full || col + j < p.N) ? __ldg(p.col_inv + col + j) : 1.f) : cinv1;
                v = (v * fminf(rinv, cinv)) * fmaxf(rinv, cinv);
              }
              if (p.epilogue == SK_EPI_BIAS || p.epilogue == SK_EPI_BIAS_RELU) {
                if (full || col + j < p.N) v += __ldg(p.bias + col + j);
              }
              if (p.epilogue == SK_EPI_BIAS_RELU || p.epilogue == SK_EPI_RELU) v = fmaxf(v, 0.f);
              accum[c0 + j] = v;
            }
            if (p.accumulate) {   // C + A @ B: the partial adjoint already in C (autodiff.pyx:30-41)
              if (full && vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 o = *reinterpret_cast<const float4 *>(crow + col + j);
                  accum[c0 + j] = __fadd_rn(o.x, accum[c0 + j]); accum[c0 + j + 1] = __fadd_rn(o.y, accum[c0 + j + 1]);
                  accum[c0 + j + 2] = __fadd_rn(o.z, accum[c0 + j + 2]); accum[c0 + j + 3] = __fadd_rn(o.w, accum[c0 + j + 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col + j < p.N) accum[c0 + j] = __fadd_rn(crow[col + j], accum[c0 + j]);
              }
            }
            if (full && vec_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4 *>(crow + col + j) =
                    make_float4(accum[c0 + j], accum[c0 + j + 1], accum[c0 + j + 2], accum[c0 + j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < p.N) crow[col + j] = accum[c0 + j];
            }
          }
        }
      }
    }
  }

  // neither CTA may exit (or free TMEM) while its partner can still touch its smem / TMEM
  tc_fence_before();
  cluster_sync_all();
  if (dyn && rank == 0 && threadIdx.x == 0) {
    // the last pair to finish re-arms the counters for the next launch on this stream
    __threadfence();
    if (atomicAdd(p.sched + 1, 1) == num_pairs - 1) {
      p.sched[0] = 0;
      p.sched[1] = 0;
      __threadfence();
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int KIND, int STAGES, int CHUNK_KB>
static int launch_kind2(const GemmProblem &g, const Operand &oa, const Operand &ob, const void *alo,
                        int64_t ld_alo, const void *blo, int64_t ld_blo, const float *row_inv = nullptr,
                        const float *col_inv = nullptr) {
  constexpr bool BF16 = KIND == KIND_BF16 || KIND == KIND_F16X3;
  constexpr bool X3 = KIND == KIND_TF32X3 || KIND == KIND_F16X3;
  constexpr int ES = BF16 ? 2 : 4;
  constexpr int BK = 128 / ES, SLAB = 128 / ES;
  constexpr size_t STAGE_BYTES = (size_t)(X3 ? 2 : 1) * (BMH * 128 + (BN2 / 2) * 128);
  constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  CUtensorMap ma, malo, mb, mblo;
  int rc;
  auto mk = [&](CUtensorMap *m, const void *base, int64_t mn, int64_t ld, bool mn_major, int tile_mn) -> int {
    if (mn_major) return make_map(m, base, ES, mn, g.K, ld, SLAB, BK, !BF16);
    return make_map(m, base, ES, g.K, mn, ld, BK, tile_mn, false);
  };
  if ((rc = mk(&ma, g.a, g.M, oa.ld, oa.mn_major, BMH))) return rc;
  if ((rc = mk(&mb, g.b, g.N, ob.ld, ob.mn_major, BN2 / 2))) return rc;
  if (X3) {
    if ((rc = mk(&malo, alo, g.M, ld_alo, oa.mn_major, BMH))) return rc;
    if ((rc = mk(&mblo, blo, g.N, ld_blo, ob.mn_major, BN2 / 2))) return rc;
  } else {
    malo = ma;
    mblo = mb;
  }
  TcParams p;
  p.c = g.c; p.bias = g.bias; p.ldc = g.ldc;
  p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
  p.epilogue = g.epilogue;
  p.row_inv = row_inv; p.col_inv = col_inv;
  p.a_inv1 = g.a_inv1; p.b_inv1 = g.b_inv1; p.accumulate = g.accumulate;
  p.tiles_m = (int)((g.M + BM2 - 1) / BM2);
  p.tiles_n = (int)((g.N + BN2 - 1) / BN2);
  static const int group_env = getenv("SOKET_B200_GEMM_GROUP_M") ? atoi(getenv("SOKET_B200_GEMM_GROUP_M")) : 8;
  p.group_m = group_env < 1 ? 1 : group_env;
  static const int dyn_env = getenv("SOKET_B200_GEMM_DYNAMIC") ? atoi(getenv("SOKET_B200_GEMM_DYNAMIC")) : 1;
  static int *sched_dev = nullptr;
  if (dyn_env && !sched_dev) {
    SK_CUDA(cudaMalloc((void **)&sched_dev, 2 * sizeof(int)));
    SK_CUDA(cudaMemsetAsync(sched_dev, 0, 2 * sizeof(int), stream()));
  }
  p.sched = dyn_env ? sched_dev : nullptr;
  const int tiles = p.tiles_m * p.tiles_n;
  // SOKET_B200_GEMM_RESERVE_SMS: SMs the persistent grid leaves alone (data-parallel training: the collective's
  // CTAs then never queue behind a GEMM CTA that holds its SM for a whole tile)
  static const int reserve_env = getenv("SOKET_B200_GEMM_RESERVE_SMS") ? atoi(getenv("SOKET_B200_GEMM_RESERVE_SMS")) : 0;
  int max_pairs = (ctx().num_sms - reserve_env) / 2;
  if (max_pairs < 1) max_pairs = 1;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
#define LAUNCH2(AMN, BMN)                                                                                  \
  do {                                                                                                     \
    auto kern = gemm_tc2_kernel<KIND, STAGES, CHUNK_KB, AMN, BMN>;                                         \
    static bool attr_set = false;                                                                          \
    if (!attr_set) {                                                                                       \
      SK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));         \
      attr_set = true;                                                                                     \
    }                                                                                                      \
    kern<<<2 * pairs, kTc2Threads, SMEM, stream()>>>(ma, malo, mb, mblo, p);                               \
  } while (0)
  if (oa.mn_major && ob.mn_major) LAUNCH2(true, true);
  else if (oa.mn_major) LAUNCH2(true, false);
  else if (ob.mn_major) LAUNCH2(false, true);
  else LAUNCH2(false, false);
#undef LAUNCH2
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int launch_gemm_tc2(const GemmProblem &g, int kind, const Operand &oa, const Operand &ob, const void *alo,
                    int64_t ld_alo, const void *blo, int64_t ld_blo, const float *row_inv, const float *col_inv) {
  if (kind == KIND_F16X3) {
    static const int chunk = getenv("SOKET_B200_F16X3_CHUNK") ? atoi(getenv("SOKET_B200_F16X3_CHUNK")) : F16X3_CHUNK_KB;
    if (chunk == 1) return launch_kind2<KIND_F16X3, 3, 1>(g, oa, ob, alo, ld_alo, blo, ld_blo, row_inv, col_inv);
    if (chunk == 2) return launch_kind2<KIND_F16X3, 3, 2>(g, oa, ob, alo, ld_alo, blo, ld_blo, row_inv, col_inv);
    return launch_kind2<KIND_F16X3, 3, 4>(g, oa, ob, alo, ld_alo, blo, ld_blo, row_inv, col_inv);
  }
  if (kind == KIND_BF16) return launch_kind2<KIND_BF16, 6, 8>(g, oa, ob, nullptr, 0, nullptr, 0);
  if (kind == KIND_TF32) return launch_kind2<KIND_TF32, 6, 1 << 20>(g, oa, ob, nullptr, 0, nullptr, 0);
  return launch_kind2<KIND_TF32X3, 3, 4>(g, oa, ob, alo, ld_alo, blo, ld_blo);
}

}  // namespace sk
