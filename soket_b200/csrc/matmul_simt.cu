// matmul_simt.cu -- CUDA-core fp32 GEMM (FFMA, exact fp32 products).
//
// The always-correct matmul: any M/N/K, any 2-D strides (so `.T` views from
// soket/tensor/ops/backward.pyx:722,734 are consumed in place), one collapsed
// batch dim.  It serves (a) shapes the tcgen05 path does not take (tiny /
// misaligned, e.g. N = 10 classes), and (b) as the on-device cross-check of the
// tensor-core kernels.  128x128x16 block tile, 8x8 register tile per thread,
// double-buffered shared memory.
#include <stdlib.h>
#include "common.cuh"
#include "matmul.cuh"

namespace sk {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;

// A_KFAST: A has unit (or small) stride along k -> load with k fastest across threads.
// B_NFAST: B has unit stride along n.
template <bool A_KFAST, bool B_NFAST>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const MMArgs p) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t tiles_n = (p.N + BN - 1) / BN;
  const int64_t tiles_m = (p.M + BM - 1) / BM;
  const int64_t tiles = tiles_m * tiles_n;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs

  for (int64_t t = blockIdx.x; t < tiles * p.batch; t += gridDim.x) {
    const int64_t bz = t / tiles;
    const int64_t tt = t - bz * tiles;
    const int64_t m0 = (tt / tiles_n) * BM, n0 = (tt % tiles_n) * BN;
    const float *A = p.a + bz * p.sa_b;
    const float *B = p.b + bz * p.sb_b;
    float *C = p.c + bz * p.sc_b;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto g_load = [&](int64_t k0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int idx = tid + e * NT;  // 0..2047 over the 128x16 tile
        int m, k;
        if (A_KFAST) { k = idx & 15; m = idx >> 4; } else { m = idx & 127; k = idx >> 7; }
        int64_t gm = m0 + m, gk = k0 + k;
        ra[e] = (gm < p.M && gk < p.K) ? A[gm * p.sa_m + gk * p.sa_k] : 0.f;
        int n, kb;
        if (B_NFAST) { n = idx & 127; kb = idx >> 7; } else { kb = idx & 15; n = idx >> 4; }
        int64_t gn = n0 + n, gkb = k0 + kb;
        rb[e] = (gn < p.N && gkb < p.K) ? B[gkb * p.sb_k + gn * p.sb_n] : 0.f;
      }
    };
    auto s_store = [&](int buf) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int idx = tid + e * NT;
        int m, k;
        if (A_KFAST) { k = idx & 15; m = idx >> 4; } else { m = idx & 127; k = idx >> 7; }
        As[buf][k][m] = ra[e];
        int n, kb;
        if (B_NFAST) { n = idx & 127; kb = idx >> 7; } else { kb = idx & 15; n = idx >> 4; }
        Bs[buf][kb][n] = rb[e];
      }
    };

    g_load(0);
    s_store(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = 0; k0 < p.K; k0 += BK) {
      const bool more = k0 + BK < p.K;
      if (more) g_load(k0 + BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float av[TM], bv[TN];
        const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
        av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) {
        s_store(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
    __syncthreads();  // smem is reused by the next tile of this block

#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (gm >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int64_t gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (gn >= p.N) continue;
        float v = acc[i][j];
        if (p.epilogue == SK_EPI_BIAS || p.epilogue == SK_EPI_BIAS_RELU) v += p.bias[gn];
        if (p.epilogue == SK_EPI_BIAS_RELU || p.epilogue == SK_EPI_RELU) v = fmaxf(v, 0.f);
        C[gm * p.ldc + gn] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ small problems
// The whole config-1 model (batch 100, hidden 100: forward.pyx:172-178 on 100 x 784 x 100
// and smaller) is ONE 128 x 128 tile for the kernel above -- a single CTA walking K
// serially (measured 177 us).  Here a CTA owns a 16 x 16 output tile and splits K over 16
// thread groups (4 x 4 threads x 4 x 4 outputs each); the 16 partial sums are added in a
// fixed order through shared memory, so the result is deterministic.
constexpr int ST = 16, SKC = 64, SKL = 16;

template <bool A_KFAST, bool B_NFAST>
__global__ void __launch_bounds__(256) gemm_small_kernel(const MMArgs p) {
  __shared__ __align__(16) float As[SKC][ST + 4];
  __shared__ __align__(16) float Bs[SKC][ST + 4];
  __shared__ __align__(16) float red[SKL][ST * ST];
  const int tid = threadIdx.x;
  const int kl = tid >> 4;            // K lane: owns k = 4*kl .. 4*kl+3 of every 64-wide chunk
  const int ot = tid & 15;            // output thread: 4 x 4 outputs at (4*oy, 4*ox)
  const int oy = ot >> 2, ox = ot & 3;
  const int64_t tiles_n = (p.N + ST - 1) / ST, tiles_m = (p.M + ST - 1) / ST;
  const int64_t tiles = tiles_m * tiles_n;
  for (int64_t t = blockIdx.x; t < tiles * p.batch; t += gridDim.x) {
    const int64_t bz = t / tiles, tt = t - bz * tiles;
    const int64_t m0 = (tt / tiles_n) * ST, n0 = (tt % tiles_n) * ST;
    const float *A = p.a + bz * p.sa_b;
    const float *B = p.b + bz * p.sb_b;
    float *C = p.c + bz * p.sc_b;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t k0 = 0; k0 < p.K; k0 += SKC) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256;   // 0..1023 over the 16 x 64 tile
        int m, k;
        if (A_KFAST) { k = idx & 63; m = idx >> 6; } else { m = idx & 15; k = idx >> 4; }
        const int64_t gm = m0 + m, gk = k0 + k;
        As[k][m] = (gm < p.M && gk < p.K) ? A[gm * p.sa_m + gk * p.sa_k] : 0.f;
        int n, kb;
        if (B_NFAST) { n = idx & 15; kb = idx >> 4; } else { kb = idx & 63; n = idx >> 6; }
        const int64_t gn = n0 + n, gkb = k0 + kb;
        Bs[kb][n] = (gn < p.N && gkb < p.K) ? B[gkb * p.sb_k + gn * p.sb_n] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = kl * 4 + kk;
        const float4 a = *reinterpret_cast<const float4 *>(&As[k][oy * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][ox * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[kl][(oy * 4 + i) * ST + ox * 4 + j] = acc[i][j];
    __syncthreads();
    {
      float v = red[0][tid];
#pragma unroll
      for (int l = 1; l < SKL; ++l) v += red[l][tid];
      const int64_t gm = m0 + (tid >> 4), gn = n0 + (tid & 15);
      if (gm < p.M && gn < p.N) {
        if (p.epilogue == SK_EPI_BIAS || p.epilogue == SK_EPI_BIAS_RELU) v += p.bias[gn];
        if (p.epilogue == SK_EPI_BIAS_RELU || p.epilogue == SK_EPI_RELU) v = fmaxf(v, 0.f);
        C[gm * p.ldc + gn] = v;
      }
    }
    __syncthreads();
  }
}

int launch_gemm_simt(const MMArgs &p) {
  if (p.M == 0 || p.N == 0 || p.batch == 0) return SK_OK;
  const bool a_kfast0 = (p.sa_k == 1) || (p.sa_m != 1);
  const bool b_nfast0 = (p.sb_n == 1) || (p.sb_k != 1);
  const int64_t big_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.batch;
  if (big_tiles < 32) {   // the 128 x 128 kernel would leave most of the 148 SMs idle
    const int64_t tiles = ((p.M + ST - 1) / ST) * ((p.N + ST - 1) / ST) * p.batch;
    const int64_t cap = (int64_t)ctx().num_sms * 8;
    const int grid = (int)(tiles < cap ? tiles : cap);
    ProfScope ps(SK_PROF_GEMM_SIMT, 2.0 * (double)p.M * (double)p.N * (double)p.K * (double)p.batch);
    if (a_kfast0 && b_nfast0) gemm_small_kernel<true, true><<<grid, 256, 0, stream()>>>(p);
    else if (a_kfast0) gemm_small_kernel<true, false><<<grid, 256, 0, stream()>>>(p);
    else if (b_nfast0) gemm_small_kernel<false, true><<<grid, 256, 0, stream()>>>(p);
    else gemm_small_kernel<false, false><<<grid, 256, 0, stream()>>>(p);
    SK_LAUNCH_CHECK();
    return SK_OK;
  }
  const int64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.batch;
  const int64_t cap = (int64_t)ctx().num_sms * 2;
  const int grid = (int)(tiles < cap ? tiles : cap);
  ProfScope ps(SK_PROF_GEMM_SIMT, 2.0 * (double)p.M * (double)p.N * (double)p.K * (double)p.batch);
  const bool a_kfast = (p.sa_k == 1) || (p.sa_m != 1);
  const bool b_nfast = (p.sb_n == 1) || (p.sb_k != 1);
  if (a_kfast && b_nfast) gemm_simt_kernel<true, true><<<grid, NT, 0, stream()>>>(p);
  else if (a_kfast) gemm_simt_kernel<true, false><<<grid, NT, 0, stream()>>>(p);
  else if (b_nfast) gemm_simt_kernel<false, true><<<grid, NT, 0, stream()>>>(p);
  else gemm_simt_kernel<false, false><<<grid, NT, 0, stream()>>>(p);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

// ------------------------------------------------------------------ float64
// np.matmul on float64 (and, through the host's cast, integer) Tensors: off the training path, so a
// plain 16 x 16 shared-memory tiled DFMA kernel; any strides, one collapsed batch dimension.
__global__ void __launch_bounds__(256) gemm_f64_kernel(const double *__restrict__ a, const double *__restrict__ b,
                                                       double *__restrict__ c, int64_t M, int64_t N, int64_t K,
                                                       int64_t sa_m, int64_t sa_k, int64_t sb_k, int64_t sb_n,
                                                       int64_t ldc, int64_t sa_b, int64_t sb_b, int64_t sc_b) {
  __shared__ double As[16][17], Bs[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t bz = blockIdx.z;
  const double *A = a + bz * sa_b;
  const double *B = b + bz * sb_b;
  const int64_t m = (int64_t)blockIdx.y * 16 + ty, n = (int64_t)blockIdx.x * 16 + tx;
  double acc = 0.0;
  for (int64_t k0 = 0; k0 < K; k0 += 16) {
    As[ty][tx] = (m < M && k0 + tx < K) ? A[m * sa_m + (k0 + tx) * sa_k] : 0.0;
    Bs[ty][tx] = (k0 + ty < K && n < N) ? B[(k0 + ty) * sb_k + n * sb_n] : 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fma(As[ty][k], Bs[k][tx], acc);
    __syncthreads();
  }
  if (m < M && n < N) c[bz * sc_b + m * ldc + n] = acc;
}

int launch_gemm_f64(const GemmProblem &g) {
  if (g.M == 0 || g.N == 0 || g.batch == 0) return SK_OK;
  SK_REQUIRE(g.batch <= 65535 && (g.M + 15) / 16 <= 65535, "matmul(float64): problem too large for the DFMA kernel");
  dim3 grid((unsigned)((g.N + 15) / 16), (unsigned)((g.M + 15) / 16), (unsigned)g.batch);
  ProfScope ps(SK_PROF_GEMM_SIMT, 2.0 * (double)g.M * (double)g.N * (double)g.K * (double)g.batch);
  gemm_f64_kernel<<<grid, 256, 0, stream()>>>((const double *)g.a, (const double *)g.b, (double *)g.c, g.M, g.N, g.K,
                                              g.sa_m, g.sa_k, g.sb_k, g.sb_n, g.ldc, g.sa_b, g.sb_b, g.sc_b);
  SK_LAUNCH_CHECK();
  return SK_OK;
}

}  // namespace sk
