// matmul.cuh -- shared GEMM problem description (SIMT + tcgen05 paths).
#pragma once
#include "common.cuh"

namespace sk {

// C[b] (M,N) = epilogue(A[b] (M,K) @ B[b] (K,N)); element strides.
struct GemmProblem {
  const void *a;
  const void *b;
  float *c;
  const float *bias;
  int a_dtype, b_dtype;
  int64_t M, N, K;
  int64_t sa_m, sa_k, sb_k, sb_n, ldc;
  int64_t batch, sa_b, sb_b, sc_b;
  int epilogue;
  // fp16x3 on pre-split operands (sk_gemm_f16x3): ONE inverse scale per operand (device scalars)
  const float *a_inv1, *b_inv1;
  int accumulate;   // C += A @ B (read-add in the epilogue) instead of C = A @ B
};

struct MMArgs {
  const float *a, *b;
  float *c;
  const float *bias;
  int64_t M, N, K;
  int64_t sa_m, sa_k, sb_k, sb_n, ldc;
  int64_t batch, sa_b, sb_b, sc_b;
  int epilogue;
};

int launch_gemm_simt(const MMArgs &p);
// float64 operands and result (g.a, g.b, g.c point at doubles; strides in elements)
int launch_gemm_f64(const GemmProblem &g);

// tcgen05 path (matmul_tc.cu)
bool tc_supported(const GemmProblem &g, int algo);
// one operand: `mn` x `k` elements of `es` bytes with element strides (s_mn, s_k)
bool tc_operand_ok(int64_t mn, int64_t k, int64_t s_mn, int64_t s_k, int es, const void *ptr);
bool tc_profitable(const GemmProblem &g);
int launch_gemm_tc(const GemmProblem &g, int algo);
bool gemm_f16x3_shape_ok(int64_t M, int64_t N, int64_t K);
int gemm_f16x3_presplit(const sk_split_operand *a, const sk_split_operand *b, float *c, int64_t ldc, int64_t M,
                        int64_t N, int64_t K, const float *bias, int epilogue, int accumulate);
int linear_bwd_f16x3(const float *adj, const float *x, const float *w, float *dx, float *dw, float *db, int64_t Bn,
                     int64_t I, int64_t O, bool *done);

}  // namespace sk
