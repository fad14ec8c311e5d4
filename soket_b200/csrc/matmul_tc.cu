// matmul_tc.cu -- tcgen05 / TMEM / TMA GEMM for sm_100a.
//
// C (M,N) fp32 = epilogue(A (M,K) @ B (K,N)), operands fp32 (kind::tf32, optionally the
// error-compensated 3xTF32 scheme) or bf16 (kind::f16), each either K-major or
// MN-major in global memory -- i.e. row-major matrices AND their `.T` views are
// consumed in place (soket/tensor/ops/forward.pyx:172-178, backward.pyx:720-736).
//
// Structure (one persistent CTA per SM, 64 + 128 * BN/128 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (128B swizzle) into a STAGES-deep
//               shared-memory ring, completion on `full` mbarriers
//   warp 1      TMEM allocator + MMA issuer: one thread issues tcgen05.mma
//               (cta_group::1, M=128, N=BN, K=32 bytes per instruction) reading A/B
//               through shared-memory matrix descriptors, fp32 accumulators in TMEM
//               (2 x BN columns, double buffered); tcgen05.commit releases ring slots
//               and publishes finished accumulators
//   warps 2..   epilogue: tcgen05.ld 32x32b.x32 -> fp32 register accumulators ->
//               (+bias)(ReLU) -> global
//
// Accumulation: the tensor core adds into the fp32 TMEM accumulator with truncation, a
// bias of ~2^-25 of the running sum per tcgen05.mma that grows LINEARLY with K (measured:
// 6e-5 relative at K = 8192).  The parity kinds therefore run the K loop in chunks of
// CHUNK_KB stages; each chunk accumulates in its own TMEM buffer (double buffered) and
// the epilogue warps drain it into round-to-nearest fp32 REGISTER accumulators
// ("promotion"), which caps the truncation bias at the chunk length (~1.5e-6).
//
// 3xTF32 (fp32 parity at 1e-5): kind::tf32 truncates operands to 10 mantissa bits
// (~1e-3).  With hi = tf32(x) taken by the hardware from the raw fp32 bits and
// lo = x - hi (exact, prepared by a small elementwise pass), A@B ~= Ahi@Bhi + Ahi@Blo +
// Alo@Bhi accumulated in fp32 in TMEM restores ~2^-21 relative accuracy at 1/3 of the
// TF32 rate.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "matmul.cuh"
#include "matmul_tc.cuh"
#include "matmul_split.cuh"

namespace sk {

// ------------------------------------------------------------------ kernel
constexpr int BM = 128;
// warps: 0 = TMA, 1 = MMA, then 4 epilogue warps per 128 accumulator columns
constexpr int tc_threads(int bn) { return 64 + 128 * (bn / 128); }

template <int KIND, int BN, int STAGES, int CHUNK_KB, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(tc_threads(BN), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_alo,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_blo,
               const TcParams p) {
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr bool X3 = KIND == KIND_TF32X3;
  constexpr int ES = BF16 ? 2 : 4;            // operand element size
  constexpr int BK = 128 / ES;                // elements per 128-byte swizzle row = K per stage
  constexpr int UK = 32 / ES;                 // K per tcgen05.mma
  constexpr int SLAB = 128 / ES;              // MN elements per 128-byte row (MN-major operands)
  constexpr uint32_t A_BYTES = BM * 128;      // BM x BK elements
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = (X3 ? 2 : 1) * (A_BYTES + B_BYTES);
  constexpr uint32_t A_LO_OFF = A_BYTES;
  constexpr uint32_t B_OFF = (X3 ? 2 : 1) * A_BYTES;
  constexpr uint32_t B_LO_OFF = B_OFF + B_BYTES;
  // descriptor strides.  K-major: 8-row groups 1024 B apart.  MN-major: 128-byte rows run
  // along MN, 8 K-rows per swizzle atom (1024 B), MN slabs BK*128 B apart.
  // fp32 MN-major: 4-row (512 B) atoms of the 32B-atom swizzle, descriptor layout type 1.
  constexpr uint32_t A_LBO = A_MN ? BK * 128 : 0, A_SBO = (A_MN && !BF16) ? 512 : 1024;
  constexpr uint32_t B_LBO = B_MN ? BK * 128 : 0, B_SBO = (B_MN && !BF16) ? 512 : 1024;
  constexpr uint32_t A_LT = (A_MN && !BF16) ? 1 : 2, B_LT = (B_MN && !BF16) ? 1 : 2;
  constexpr uint32_t A_KSTEP = A_MN ? UK * 128 : 32;   // bytes to advance per UK along K
  constexpr uint32_t B_KSTEP = B_MN ? UK * 128 : 32;
  constexpr uint32_t IDESC = make_idesc(BF16 ? 1 : 2, A_MN, B_MN, BM, BN);
  constexpr int TMEM_COLS = 2 * BN;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle (address bits [7,10) select the XOR)
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t *full_bar = bars;                    // [STAGES]
  uint64_t *empty_bar = bars + STAGES;          // [STAGES]
  uint64_t *tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t *tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (X3) { tma_prefetch_desc(&map_alo); tma_prefetch_desc(&map_blo); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4 * (BN / 128));   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % p.tiles_m) * BM, n0 = (tile / p.tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          mbar_expect_tx(fb, STAGE_BYTES);
          const int k0 = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / SLAB; ++j) {
              tma_load_2d(sbase + j * (BK * 128), &map_a, fb, m0 + j * SLAB, k0);
              if (X3) tma_load_2d(sbase + A_LO_OFF + j * (BK * 128), &map_alo, fb, m0 + j * SLAB, k0);
            }
          } else {
            tma_load_2d(sbase, &map_a, fb, k0, m0);
            if (X3) tma_load_2d(sbase + A_LO_OFF, &map_alo, fb, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / SLAB; ++j) {
              tma_load_2d(sbase + B_OFF + j * (BK * 128), &map_b, fb, n0 + j * SLAB, k0);
              if (X3) tma_load_2d(sbase + B_LO_OFF + j * (BK * 128), &map_blo, fb, n0 + j * SLAB, k0);
            }
          } else {
            tma_load_2d(sbase + B_OFF, &map_b, fb, k0, n0);
            if (X3) tma_load_2d(sbase + B_LO_OFF, &map_blo, fb, k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB) {
          const int kb1 = kb0 + CHUNK_KB < num_kb ? kb0 + CHUNK_KB : num_kb;
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);   // epilogue drained this buffer
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);          // TMA bytes have landed
            tc_fence_after();
            const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da = make_smem_desc(sbase + k * A_KSTEP, A_LBO, A_SBO, A_LT);
              const uint64_t db = make_smem_desc(sbase + B_OFF + k * B_KSTEP, B_LBO, B_SBO, B_LT);
              tc_mma<BF16>(d_tmem, da, db, IDESC, ((kb - kb0) | k) != 0);
              if (X3) {
                const uint64_t dalo = make_smem_desc(sbase + A_LO_OFF + k * A_KSTEP, A_LBO, A_SBO, A_LT);
                const uint64_t dblo = make_smem_desc(sbase + B_LO_OFF + k * B_KSTEP, B_LBO, B_SBO, B_LT);
                tc_mma<BF16>(d_tmem, da, dblo, IDESC, 1);     // Ahi @ Blo
                tc_mma<BF16>(d_tmem, dalo, db, IDESC, 1);     // Alo @ Bhi
              }
            }
            tc_commit(smem_u32(&empty_bar[stage]));              // slot free once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(smem_u32(&tfull_bar[acc]));                  // this chunk's partial sums are complete
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..) =====================
    // warp w owns TMEM lanes 32*(w%4).. (hardware rule) and accumulator columns
    // 128*((w-2)/4)..+128; each thread carries one output row x 128 columns in registers.
    const int q = warp & 3;
    const int cbase = ((warp - 2) >> 2) * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_ok = (p.ldc % 4 == 0) && ((((uintptr_t)p.c) & 15) == 0);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % p.tiles_m) * BM, n0 = (tile / p.tiles_m) * BN + cbase;
      float accum[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) accum[j] = 0.f;
      const bool single = num_kb <= CHUNK_KB;
      for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB) {
        mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
        tc_fence_after();
        if (n0 < p.N) {                      // warp-uniform: this column block exists
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + cbase + c0), r);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              accum[c0 + j] = single ? __uint_as_float(r[j]) : __fadd_rn(accum[c0 + j], __uint_as_float(r[j]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      const int row = m0 + q * 32 + lane;
      if (row < p.M && n0 < p.N) {
        float *crow = p.c + (int64_t)row * p.ldc;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          const int col = n0 + c0;
          if (col < p.N) {
            const bool full = col + 32 <= p.N;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = accum[c0 + j];
              if (p.epilogue == SK_EPI_BIAS || p.epilogue == SK_EPI_BIAS_RELU) {
                if (full || col + j < p.N) v += __ldg(p.bias + col + j);
              }
              if (p.epilogue == SK_EPI_BIAS_RELU || p.epilogue == SK_EPI_RELU) v = fmaxf(v, 0.f);
              accum[c0 + j] = v;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// lo = tf32_rn(x - tf32_trunc(x)) for a (rows, cols) matrix with unit column stride.
// The hardware takes hi = tf32_trunc(x) from the raw fp32 bits, so x - hi is exact; the
// tensor core would TRUNCATE lo to 10 mantissa bits too, and a truncation always errs
// towards zero -- a bias of ~1e-6 * |a||b| that does not average out over K.  Rounding
// lo to nearest here makes the residual unbiased (it then shrinks like 1/sqrt(K)).
__device__ __forceinline__ float split_lo(float v) {
  const float lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
  return __uint_as_float(r);
}

__global__ void __launch_bounds__(256)
split_lo_kernel(const float *__restrict__ x, int64_t ldx, float *__restrict__ lo, int64_t ldlo,
                int64_t rows, int64_t cols4) {
  const int64_t total = rows * cols4;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / cols4, c = (i - r * cols4) << 2;
    const float4 v = ld_stream(reinterpret_cast<const float4 *>(x + r * ldx + c));
    float4 o;
    o.x = split_lo(v.x); o.y = split_lo(v.y); o.z = split_lo(v.z); o.w = split_lo(v.w);
    *reinterpret_cast<float4 *>(lo + r * ldlo + c) = o;   // re-read soon by TMA: keep in L2
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D tensor map over a matrix stored as `outer` rows of `inner` contiguous elements
// (row pitch ld elements), box = (box_inner x box_outer), 128-byte swizzle (16-byte atoms,
// or 32-byte atoms for MN-major fp32 operands), zero OOB fill.
int make_map(CUtensorMap *m, const void *base, int es, int64_t inner, int64_t outer, int64_t ld,
                    int box_inner, int box_outer, bool atom32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("matmul(tcgen05): cuTensorMapEncodeTiled is not available from the driver");
    return SK_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("matmul(tcgen05): cuTensorMapEncodeTiled failed with %d (inner=%lld outer=%lld ld=%lld box=%dx%d)",
              (int)r, (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer);
    return SK_ERR_CUDA;
  }
  return SK_OK;
}

static bool classify(int64_t mn, int64_t k, int64_t s_mn, int64_t s_k, int es, const void *ptr, Operand &o) {
  if ((((uintptr_t)ptr) & 15) != 0) return false;
  const int64_t align = 16 / es;   // TMA: base and pitch must be multiples of 16 bytes
  if (s_k == 1 || k == 1) {        // K-major: `mn` rows of k contiguous elements, pitch s_mn
    const int64_t ld = mn > 1 ? s_mn : (k + align - 1) / align * align;
    if (ld % align == 0 && ld >= k) {
      o.mn_major = false;
      o.ld = ld;
      return true;
    }
  }
  if (s_mn == 1 || mn == 1) {      // MN-major: k rows of `mn` contiguous elements, pitch s_k
    const int64_t ld = k > 1 ? s_k : (mn + align - 1) / align * align;
    if (ld % align == 0 && ld >= mn) {
      o.mn_major = true;
      o.ld = ld;
      return true;
    }
  }
  return false;
}

bool tc_operand_ok(int64_t mn, int64_t k, int64_t s_mn, int64_t s_k, int es, const void *ptr) {
  Operand o;
  return classify(mn, k, s_mn, s_k, es, ptr, o);
}

bool tc_pairable(const GemmProblem &g) {
  static const int want_2cta = getenv("SOKET_B200_GEMM_2CTA") ? atoi(getenv("SOKET_B200_GEMM_2CTA")) : 1;
  return want_2cta && g.M >= 256 && g.N >= 128;
}

bool tc_supported(const GemmProblem &g, int algo) {
  if (algo == SK_MM_F16X3 && !(tc_pairable(g) && g.K >= 64)) return false;   // CTA-pair kernel only
  const bool bf16 = algo == SK_MM_BF16;
  if (bf16 != (g.a_dtype == SK_BF16)) return false;
  if (g.M < 1 || g.N < 1 || g.K < 1) return false;
  if (g.M > INT32_MAX || g.N > INT32_MAX || g.K > INT32_MAX) return false;
  Operand a, b;
  const int es = bf16 ? 2 : 4;
  if (!classify(g.M, g.K, g.sa_m, g.sa_k, es, g.a, a)) return false;
  if (!classify(g.N, g.K, g.sb_n, g.sb_k, es, g.b, b)) return false;
  return encode_fn() != nullptr;
}

bool tc_profitable(const GemmProblem &g) {
  // tiny problems (the whole C1 model) stay on the FFMA kernel: one CTA wave of tcgen05
  // setup costs more than they take.  Skinny ones (N = 10 classifier) do go to tcgen05.
  return g.K >= 8 && (g.M >= 64 || g.N >= 64) && (double)g.M * g.N * g.K >= 4.0 * 128 * 128 * 128;
}

template <int KIND, int BN, int STAGES, int CHUNK_KB>
static int launch_kind(const GemmProblem &g, const Operand &oa, const Operand &ob, const float *alo,
                       int64_t ld_alo, const float *blo, int64_t ld_blo) {
  constexpr bool BF16 = KIND == KIND_BF16;
  constexpr bool X3 = KIND == KIND_TF32X3;
  constexpr int ES = BF16 ? 2 : 4;
  constexpr int BK = 128 / ES, SLAB = 128 / ES;
  constexpr size_t STAGE_BYTES = (size_t)(X3 ? 2 : 1) * (BM * 128 + BN * 128);
  constexpr size_t SMEM = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  CUtensorMap ma, malo, mb, mblo;
  int rc;
  auto mk = [&](CUtensorMap *m, const void *base, int64_t mn, int64_t ld, bool mn_major, int tile_mn) -> int {
    if (mn_major) return make_map(m, base, ES, mn, g.K, ld, SLAB, BK, !BF16);
    return make_map(m, base, ES, g.K, mn, ld, BK, tile_mn, false);
  };
  if ((rc = mk(&ma, g.a, g.M, oa.ld, oa.mn_major, BM))) return rc;
  if ((rc = mk(&mb, g.b, g.N, ob.ld, ob.mn_major, BN))) return rc;
  if (X3) {
    if ((rc = mk(&malo, alo, g.M, ld_alo, oa.mn_major, BM))) return rc;
    if ((rc = mk(&mblo, blo, g.N, ld_blo, ob.mn_major, BN))) return rc;
  } else {
    malo = ma;
    mblo = mb;
  }
  TcParams p;
  p.c = g.c; p.bias = g.bias; p.ldc = g.ldc;
  p.M = (int)g.M; p.N = (int)g.N; p.K = (int)g.K;
  p.epilogue = g.epilogue;
  p.row_inv = nullptr; p.col_inv = nullptr;
  p.a_inv1 = nullptr; p.b_inv1 = nullptr; p.accumulate = 0;
  p.group_m = 1;
  p.sched = nullptr;
  p.tiles_m = (int)((g.M + BM - 1) / BM);
  p.tiles_n = (int)((g.N + BN - 1) / BN);
  const int tiles = p.tiles_m * p.tiles_n;
  const int grid = tiles < ctx().num_sms ? tiles : ctx().num_sms;
#define LAUNCH(AMN, BMN)                                                                                   \
  do {                                                                                                     \
    auto kern = gemm_tc_kernel<KIND, BN, STAGES, CHUNK_KB, AMN, BMN>;                                                \
    static bool attr_set = false;                                                                          \
    if (!attr_set) {                                                                                       \
      SK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));         \
      attr_set = true;                                                                                     \
    }                                                                                                      \
    kern<<<grid, tc_threads(BN), SMEM, stream()>>>(ma, malo, mb, mblo, p);                                     \
  } while (0)
  if (oa.mn_major && ob.mn_major) LAUNCH(true, true);
  else if (oa.mn_major) LAUNCH(true, false);
  else if (ob.mn_major) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  SK_LAUNCH_CHECK();
  return SK_OK;
}

static int make_lo(const void *src, const Operand &o, int64_t mn, int64_t k, float **lo, int64_t *ld_lo) {
  // the operand as stored: `outer` rows of `inner` contiguous elements
  const int64_t inner = o.mn_major ? mn : k, outer = o.mn_major ? k : mn;
  const int64_t ld = (inner + 3) / 4 * 4;
  int rc = sk_malloc((size_t)(outer * ld) * sizeof(float), (void **)lo);
  if (rc) return rc;
  *ld_lo = ld;
  const int64_t cols4 = ld / 4;   // reads up to 3 elements past `inner` inside the source pitch (ld <= o.ld)
  int grid = grid_for(outer * cols4, 256, 8);
  split_lo_kernel<<<grid, 256, 0, stream()>>>((const float *)src, o.ld, *lo, ld, outer, cols4);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

int launch_gemm_tc(const GemmProblem &g0, int algo) {
  GemmProblem g = g0;
  Operand oa, ob;
  const int es = algo == SK_MM_BF16 ? 2 : 4;
  if (algo == SK_MM_F16X3 && !tc_supported(g, algo)) {
    set_error("matmul(fp16x3): needs M >= 256, N >= 128, K >= 64 and TMA-describable fp32 operands");
    return SK_ERR_UNSUPPORTED;
  }
  if (!classify(g.M, g.K, g.sa_m, g.sa_k, es, g.a, oa) || !classify(g.N, g.K, g.sb_n, g.sb_k, es, g.b, ob)) {
    set_error("matmul(tcgen05): unsupported operand layout");
    return SK_ERR_UNSUPPORTED;
  }
  const double flops = 2.0 * (double)g.M * (double)g.N * (double)g.K;
  for (int64_t bz = 0; bz < g.batch; ++bz) {
    GemmProblem gi = g;
    gi.a = (const char *)g.a + bz * g.sa_b * es;
    gi.b = (const char *)g.b + bz * g.sb_b * es;
    gi.c = g.c + bz * g.sc_b;
    int rc;
    const bool pair = tc_pairable(g);
    const double prep_bytes = 8.0 * ((double)g.M * (double)g.K + (double)g.N * (double)g.K);
    if (algo == SK_MM_F16X3) {
      // fp32 operands -> fp16 hi/lo pairs with one power-of-two scale per row of A / column of B
      SplitOperand sa, sb;
      const bool a_rows = !oa.mn_major, b_rows = !ob.mn_major;   // K-major: the scale index is the stored row
      {
        ProfScope pp(SK_PROF_GEMM_PREP, prep_bytes);
        if ((rc = split_f16((const float *)gi.a, oa.ld, a_rows ? g.M : g.K, a_rows ? g.K : g.M, a_rows, sa))) return rc;
        if ((rc = split_f16((const float *)gi.b, ob.ld, b_rows ? g.N : g.K, b_rows ? g.K : g.N, b_rows, sb))) {
          sa.release();
          return rc;
        }
      }
      GemmProblem gh = gi;
      gh.a = sa.hi; gh.b = sb.hi;
      Operand ha = oa, hb = ob;
      ha.ld = sa.ld; hb.ld = sb.ld;
      {
        ProfScope ps(SK_PROF_GEMM_TC, flops);
        rc = launch_gemm_tc2(gh, KIND_F16X3, ha, hb, sa.lo, sa.ld, sb.lo, sb.ld, sa.inv_scale, sb.inv_scale);
      }
      sa.release();   // stream-ordered: reusable only by later work on the same stream
      sb.release();
    } else if (algo == SK_MM_BF16) {
      ProfScope ps(SK_PROF_GEMM_TC, flops);
      rc = pair ? launch_gemm_tc2(gi, KIND_BF16, oa, ob, nullptr, 0, nullptr, 0)
                : launch_kind<KIND_BF16, 256, 4, 8>(gi, oa, ob, nullptr, 0, nullptr, 0);    // promote every K = 512
    } else if (algo == SK_MM_TF32) {
      ProfScope ps(SK_PROF_GEMM_TC, flops);
      rc = pair ? launch_gemm_tc2(gi, KIND_TF32, oa, ob, nullptr, 0, nullptr, 0)
                : launch_kind<KIND_TF32, 256, 4, 1 << 20>(gi, oa, ob, nullptr, 0, nullptr, 0);  // 1e-3 class: no promotion
    } else {
      float *alo = nullptr, *blo = nullptr;
      int64_t ld_alo = 0, ld_blo = 0;
      {
        ProfScope pp(SK_PROF_GEMM_PREP, prep_bytes);
        if ((rc = make_lo(gi.a, oa, g.M, g.K, &alo, &ld_alo))) return rc;
        if ((rc = make_lo(gi.b, ob, g.N, g.K, &blo, &ld_blo))) { sk_free(alo); return rc; }
      }
      {
        ProfScope ps(SK_PROF_GEMM_TC, flops);
        if (pair) rc = launch_gemm_tc2(gi, KIND_TF32X3, oa, ob, alo, ld_alo, blo, ld_blo);
        else if (g.N > 128) rc = launch_kind<KIND_TF32X3, 256, 2, 4>(gi, oa, ob, alo, ld_alo, blo, ld_blo);
        else rc = launch_kind<KIND_TF32X3, 128, 3, 4>(gi, oa, ob, alo, ld_alo, blo, ld_blo);  // promote every K = 128
      }
      sk_free(alo);   // stream-ordered: reusable only by later work on the same stream
      sk_free(blo);
    }
    if (rc) return rc;
  }
  return SK_OK;
}

// Linear backward on SHARED operand splits (prototypes.pyx:108-115 backward = backward.pyx:720-736):
//   dX (B, I) = adj (B, O) @ W(I, O).T        dW (I, O) = X(B, I).T @ adj (B, O)
// adj is split ONCE, by rows (scale 2^e_b per sample).  The dX GEMM consumes it as its K-major
// A operand.  For the dW GEMM the same hi/lo arrays are the MN-major B operand (K = batch);
// their per-row scale is not constant along K there, so it is folded into the other operand:
// X'[b, i] = X[b, i] * 2^-e_b (exact), split by columns.  One split pass over adj (12 B/elem
// for the separate column split) disappears per layer.  All three matrices row-major
// contiguous, pitches multiples of 16 bytes.  *done = false: shapes not eligible, nothing ran.
int linear_bwd_f16x3(const float *adj, const float *x, const float *w, float *dx, float *dw, float *db,
                     int64_t Bn, int64_t I, int64_t O, bool *done) {
  *done = false;
  static const bool fuse_colsum = !(getenv("SOKET_B200_FUSE_COLSUM") && !strcmp(getenv("SOKET_B200_FUSE_COLSUM"), "0"));
  GemmProblem gx, gw;
  memset(&gx, 0, sizeof(gx));
  memset(&gw, 0, sizeof(gw));
  gx.a = adj; gx.b = w; gx.c = dx; gx.a_dtype = gx.b_dtype = SK_F32;
  gx.M = Bn; gx.K = O; gx.N = I; gx.sa_m = O; gx.sa_k = 1; gx.sb_k = 1; gx.sb_n = O; gx.ldc = I; gx.batch = 1;
  gw.a = x; gw.b = adj; gw.c = dw; gw.a_dtype = gw.b_dtype = SK_F32;
  gw.M = I; gw.K = Bn; gw.N = O; gw.sa_m = 1; gw.sa_k = I; gw.sb_k = O; gw.sb_n = 1; gw.ldc = O; gw.batch = 1;
  if (!dx || !dw || !tc_supported(gx, SK_MM_F16X3) || !tc_supported(gw, SK_MM_F16X3)) return SK_OK;
  int rc;
  SplitOperand sa, sw, sx;
  {
    ProfScope pp(SK_PROF_GEMM_PREP, 8.0 * ((double)Bn * O + (double)I * O + (double)Bn * I));
    // db = column sums of adj (autodiff.pyx:84) ride along with the row split of adj when they can
    const bool db_here = db && fuse_colsum && split_colsum_supported(O);
    if ((rc = split_f16(adj, O, Bn, O, true, sa, nullptr, db_here ? db : nullptr))) return rc;
    if (db && !db_here && (rc = sk_colsum(adj, nullptr, db, Bn, O))) { sa.release(); return rc; }
    if ((rc = split_f16(w, O, I, O, true, sw))) { sa.release(); return rc; }                    // W rows = N index of dX
    if ((rc = split_f16(x, I, Bn, I, false, sx, sa.inv_scale))) { sa.release(); sw.release(); return rc; }
  }
  Operand k_major, mn_major;
  k_major.mn_major = false;
  mn_major.mn_major = true;
  {
    GemmProblem gh = gx;
    gh.a = sa.hi; gh.b = sw.hi;
    Operand oa = k_major, ob = k_major;
    oa.ld = sa.ld; ob.ld = sw.ld;
    ProfScope ps(SK_PROF_GEMM_TC, 2.0 * (double)Bn * I * O);
    rc = launch_gemm_tc2(gh, KIND_F16X3, oa, ob, sa.lo, sa.ld, sw.lo, sw.ld, sa.inv_scale, sw.inv_scale);
  }
  if (rc == SK_OK) {
    GemmProblem gh = gw;
    gh.a = sx.hi; gh.b = sa.hi;
    Operand oa = mn_major, ob = mn_major;
    oa.ld = sx.ld; ob.ld = sa.ld;
    ProfScope ps(SK_PROF_GEMM_TC, 2.0 * (double)Bn * I * O);
    rc = launch_gemm_tc2(gh, KIND_F16X3, oa, ob, sx.lo, sx.ld, sa.lo, sa.ld, sx.inv_scale, nullptr);
  }
  sa.release();
  sw.release();
  sx.release();
  *done = rc == SK_OK;
  return rc;
}

// fp16x3 GEMM on operands that were split BEFORE the call (sk_split_f16, or emitted by the kernel that
// produced the matrix): no operand pass here.  Each operand carries ONE power-of-two scale, so the
// same hi / lo pair serves as a K-major operand in one GEMM and as an MN-major operand in another
// (a Linear layer's input X in forward and in dW = X.T @ adj; its weight in forward and in dX).
bool gemm_f16x3_shape_ok(int64_t M, int64_t N, int64_t K) { return M >= 256 && N >= 128 && K >= 64; }

int gemm_f16x3_presplit(const sk_split_operand *a, const sk_split_operand *b, float *c, int64_t ldc, int64_t M,
                        int64_t N, int64_t K, const float *bias, int epilogue, int accumulate) {
  SK_REQUIRE(a && b && c && a->hi && a->lo && b->hi && b->lo && a->scale && b->scale, "sk_gemm_f16x3: null operand");
  SK_REQUIRE(gemm_f16x3_shape_ok(M, N, K), "sk_gemm_f16x3: needs M >= 256, N >= 128, K >= 64 (got %lld x %lld x %lld)",
             (long long)M, (long long)N, (long long)K);
  SK_REQUIRE(M <= INT32_MAX && N <= INT32_MAX && K <= INT32_MAX, "sk_gemm_f16x3: dimension too large");
  SK_REQUIRE(a->ld % 8 == 0 && b->ld % 8 == 0, "sk_gemm_f16x3: operand pitches must be multiples of 8 elements");
  SK_REQUIRE(a->ld >= (a->mn_major ? M : K) && b->ld >= (b->mn_major ? N : K), "sk_gemm_f16x3: pitch below row length");
  const uintptr_t al = (uintptr_t)a->hi | (uintptr_t)a->lo | (uintptr_t)b->hi | (uintptr_t)b->lo;
  SK_REQUIRE((al & 15) == 0, "sk_gemm_f16x3: operands must be 16-byte aligned");
  SK_REQUIRE(ldc >= N, "sk_gemm_f16x3: ldc below N");
  SK_REQUIRE(epilogue >= SK_EPI_NONE && epilogue <= SK_EPI_RELU, "sk_gemm_f16x3: bad epilogue %d", epilogue);
  SK_REQUIRE(!(epilogue == SK_EPI_BIAS || epilogue == SK_EPI_BIAS_RELU) || bias, "sk_gemm_f16x3: epilogue needs a bias");
  SK_REQUIRE(encode_fn() != nullptr, "sk_gemm_f16x3: cuTensorMapEncodeTiled is not available from the driver");
  GemmProblem g;
  memset(&g, 0, sizeof(g));
  g.a = a->hi; g.b = b->hi; g.c = c; g.bias = bias;
  g.a_dtype = g.b_dtype = SK_F32;
  g.M = M; g.N = N; g.K = K; g.ldc = ldc; g.batch = 1;
  g.epilogue = epilogue;
  g.a_inv1 = a->scale + 1; g.b_inv1 = b->scale + 1;
  g.accumulate = accumulate ? 1 : 0;
  Operand oa, ob;
  oa.mn_major = a->mn_major != 0; oa.ld = a->ld;
  ob.mn_major = b->mn_major != 0; ob.ld = b->ld;
  ProfScope ps(SK_PROF_GEMM_TC, 2.0 * (double)M * (double)N * (double)K);
  return launch_gemm_tc2(g, KIND_F16X3, oa, ob, a->lo, a->ld, b->lo, b->ld, nullptr, nullptr);
}

}  // namespace sk
