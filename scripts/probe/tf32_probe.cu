// Standalone probe: one 128x128 tile, K = 32 fp32 (one 128B-swizzled stage), tcgen05 kind::tf32
// vs kind::f16 (bf16), dumping the shared-memory tile and the accumulator.  Not part of the product.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool BF16>
__global__ void probe(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb, float *out, float *dump, uint32_t idesc, int nk) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bar = (uint64_t *)(smem + 32768);
  uint32_t *slot = (uint32_t *)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    uint32_t fb = smem_u32(&bar[0]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(32768u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)), "l"(&ma), "r"(fb), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem + 16384)), "l"(&mb), "r"(fb), "r"(0), "r"(0) : "memory");
    mbar_wait(fb, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int i = 0; i < 64; ++i) { dump[i] = ((float *)smem)[i]; dump[64 + i] = ((float *)(smem + 16384))[i]; }
    for (int k = 0; k < nk; ++k) {
      uint64_t da = make_desc(smem_u32(smem) + k * 32, 0, 1024);
      uint64_t db = make_desc(smem_u32(smem + 16384) + k * 32, 0, 1024);
      uint32_t acc = k != 0;
      if (BF16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
  }
  mbar_wait(smem_u32(&bar[1]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  for (int c0 = 0; c0 < 128; c0 += 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap mk(EncodeFn fn, void *base, CUtensorMapDataType dt, int es, int inner, int outer, int box_inner, int box_outer) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * es};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeFn fn = (EncodeFn)p;
  const int M = 128, N = 128;
  float *out, *dump; CK(cudaMalloc(&out, M * N * 4)); CK(cudaMalloc(&dump, 128 * 4));
  std::vector<float> ho(M * N), hd(128);
  // ---- fp32 / tf32: K = 32
  {
    const int K = 32;
    std::vector<float> a(M * K), b(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) a[i * K + k] = (float)((i + k) % 5 - 2);
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) b[j * K + k] = (float)((j * 3 + k) % 7 - 3);
    float *da, *db; CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    for (int variant = 0; variant < 3; ++variant) {
      CUtensorMapDataType dt = variant == 1 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
      CUtensorMap ma = mk(fn, da, dt, 4, K, M, 32, 128), mb = mk(fn, db, dt, 4, K, N, 32, 128);
      uint32_t fmt = 2;
      uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      if (variant == 2) idesc |= 0;  // placeholder for further variants
      CK(cudaMemset(out, 0xff, M * N * 4));
      probe<false><<<1, 128, 40000>>>(ma, mb, out, dump, idesc, 4);
      cudaError_t e = cudaDeviceSynchronize();
      printf("tf32 variant %d: sync=%s\n", variant, cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
      CK(cudaMemcpy(ho.data(), out, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hd.data(), dump, 128 * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int nz = 0;
      for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double r = 0; for (int k = 0; k < K; ++k) r += a[i * K + k] * b[j * K + k]; maxerr = fmax(maxerr, fabs(r - ho[i * N + j])); nz += ho[i * N + j] != 0; }
      printf("  maxerr %.4f nonzero %d  out[0,0..3]= %g %g %g %g   smemA[0..7]= %g %g %g %g %g %g %g %g  smemB[0..3]= %g %g %g %g\n", maxerr, nz, ho[0], ho[1], ho[2], ho[3], hd[0], hd[1], hd[2], hd[3], hd[4], hd[5], hd[6], hd[7], hd[64], hd[65], hd[66], hd[67]);
    }
  }
  // ---- bf16: K = 64
  {
    const int K = 64;
    std::vector<__nv_bfloat16> a(M * K), b(N * K);
    std::vector<float> af(M * K), bf(N * K);
    for (int i = 0; i < M; ++i) for (int k = 0; k < K; ++k) { af[i * K + k] = (float)((i + k) % 5 - 2); a[i * K + k] = __float2bfloat16(af[i * K + k]); }
    for (int j = 0; j < N; ++j) for (int k = 0; k < K; ++k) { bf[j * K + k] = (float)((j * 3 + k) % 7 - 3); b[j * K + k] = __float2bfloat16(bf[j * K + k]); }
    void *da, *db; CK(cudaMalloc(&da, a.size() * 2)); CK(cudaMalloc(&db, b.size() * 2));
    CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    CUtensorMap ma = mk(fn, da, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, 64, 128), mb = mk(fn, db, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, 64, 128);
    uint32_t fmt = 1;
    uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    probe<true><<<1, 128, 40000>>>(ma, mb, out, dump, idesc, 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("bf16: sync=%s\n", cudaGetErrorString(e));
    CK(cudaMemcpy(ho.data(), out, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) { double r = 0; for (int k = 0; k < K; ++k) r += af[i * K + k] * bf[j * K + k]; maxerr = fmax(maxerr, fabs(r - ho[i * N + j])); }
    printf("  maxerr %.4f out[0,0..3]= %g %g %g %g\n", maxerr, ho[0], ho[1], ho[2], ho[3]);
  }
  return 0;
}
