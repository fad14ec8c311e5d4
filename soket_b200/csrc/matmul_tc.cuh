// matmul_tc.cuh -- PTX helpers and descriptors shared by the tcgen05 GEMM kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "matmul.cuh"

namespace sk {

// ------------------------------------------------------------------ device PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (CUDA error at the next sync) instead of hanging.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  if (BF16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (tcgen05 "smem descriptor", SWIZZLE_128B):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte
//   offset >> 4   [46,48) version = 1 (Blackwell)   [61,64) layout type (2 = 128B swizzle)
//   layout type 1 = "128B swizzle with 32-byte atoms": the ONLY layout tcgen05 accepts
//   for MN-major 32-bit (tf32) operands -- 4 K-rows of 128 B per atom, the four 32 B
//   chunks of a row XOR-ed with (row % 4); TMA writes it with SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Instruction descriptor (upper 32 bits of the "runtime idesc"):
//   [4,6) D format (1 = f32)  [7,10) A format  [10,13) B format (0 f16, 1 bf16, 2 tf32)
//   [15] A major (0 K, 1 MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int fmt, bool a_mn, bool b_mn, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((a_mn ? 1u : 0u) << 15) |
         ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

enum { KIND_TF32 = 0, KIND_TF32X3 = 1, KIND_BF16 = 2, KIND_F16X3 = 3 };
// fp16x3: K blocks (64 elements each) per TMEM accumulation chunk before promotion to registers
constexpr int F16X3_CHUNK_KB = 4;

struct TcParams {
  float *c;
  const float *bias;
  int64_t ldc;
  int M, N, K;
  int epilogue;
  int tiles_m, tiles_n;
  int group_m;            // M-tiles per rasterisation group (tile_coords)
  const float *row_inv;   // fp16x3: 2^-e per output row / column (operand scales to undo)
  const float *col_inv;
  const float *a_inv1;    // fp16x3 on pre-split operands: one inverse scale for all of A / all of B
  const float *b_inv1;
  int accumulate;         // epilogue adds the result to what C holds (in-place gradient accumulation)
  int *sched;             // CTA-pair kernel: {next tile, finished pairs} for dynamic tile scheduling (NULL = static)
};

// Tile rasterisation: groups of `group_m` M-tiles are swept across all N-tiles (M fastest
// inside the group), so the tiles in flight at any moment share a few A row-blocks and a few
// B column-blocks that stay L2-resident; the plain M-fastest order re-read A from DRAM once
// per wave (ncu: 1.05 GB per 8192x4096x4096 launch against 0.2 GB of operands).
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int &tm, int &tn) {
  const int per_group = group_m * tiles_n;
  const int g = tile / per_group, within = tile - g * per_group;
  const int m_base = g * group_m;
  const int rows = tiles_m - m_base < group_m ? tiles_m - m_base : group_m;
  tn = within / rows;
  tm = m_base + (within - tn * rows);
  if (g & 1) tn = tiles_n - 1 - tn;   // serpentine: the next group starts on the B columns still in L2
}

struct Operand {
  bool mn_major;   // unit stride runs along M/N instead of K
  int64_t ld;      // pitch (elements) of the non-unit dimension
};

int make_map(CUtensorMap *m, const void *base, int es, int64_t inner, int64_t outer, int64_t ld,
             int box_inner, int box_outer, bool atom32);

// cta_group::2 kernels (matmul_tc2.cu)
int launch_gemm_tc2(const GemmProblem &g, int kind, const Operand &oa, const Operand &ob, const void *alo,
                    int64_t ld_alo, const void *blo, int64_t ld_blo, const float *row_inv = nullptr,
                    const float *col_inv = nullptr);

}  // namespace sk
