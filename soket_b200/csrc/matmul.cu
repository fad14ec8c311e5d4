// matmul.cu -- sk_matmul / sk_linear_fwd dispatch.
// Replaces _MATMUL (soket/tensor/ops/intern.pyx:63) as called from
// forward.pyx:172-178 (x @ y) and backward.pyx:704-742 (adj @ y.T, x.T @ adj),
// and the Linear(+ReLU) module sequence prototypes.pyx:108-115,302.
#include "common.cuh"
#include "matmul.cuh"
#include "matmul_split.cuh"
#include <stdlib.h>
#include <string.h>

namespace sk {

static int parse_2d(const sk_array *a, const sk_array *b, const sk_array *out, GemmProblem &g) {
  SK_REQUIRE(a && b && out, "matmul: null array");
  SK_REQUIRE(a->ndim >= 2 && b->ndim >= 2, "matmul: operands must be at least 2-D");
  SK_REQUIRE(out->ndim >= 2 && out->ndim >= a->ndim && out->ndim >= b->ndim, "matmul: bad output rank");
  g.M = a->shape[a->ndim - 2];
  g.K = a->shape[a->ndim - 1];
  g.N = b->shape[b->ndim - 1];
  SK_REQUIRE(b->shape[b->ndim - 2] == g.K, "matmul: inner dimensions differ (%lld vs %lld)",
             (long long)g.K, (long long)b->shape[b->ndim - 2]);
  SK_REQUIRE(out->shape[out->ndim - 2] == g.M && out->shape[out->ndim - 1] == g.N,
             "matmul: output shape mismatch");
  SK_REQUIRE(out->dtype == SK_F32 || (out->dtype == SK_F64 && a->dtype == SK_F64 && b->dtype == SK_F64),
             "matmul: output must be float32 (or float64 for float64 operands)");
  SK_REQUIRE(out->strides[out->ndim - 1] == 1 || g.N == 1, "matmul: output rows must be contiguous");
  g.sa_m = a->strides[a->ndim - 2];
  g.sa_k = a->strides[a->ndim - 1];
  g.sb_k = b->strides[b->ndim - 2];
  g.sb_n = b->strides[b->ndim - 1];
  g.ldc = g.M > 1 ? out->strides[out->ndim - 2] : g.N;
  return SK_OK;
}

static int run_one(const GemmProblem &g0, int algo);

// TMA needs 16-byte aligned bases and row pitches.  Operands that miss that (e.g. the
// (4096, 10) classifier weight: pitch 40 B) are copied once into a pitch-aligned scratch
// matrix -- 8 B per element against 2*M*N*K flops -- and the problem re-enters the
// tensor-core path.  The copy keeps the operand's major-ness.
static int repitch(const void *src, int64_t mn, int64_t k, int64_t s_mn, int64_t s_k, float **out,
                   int64_t *o_mn, int64_t *o_k) {
  const bool k_major = (s_k == 1 || k == 1) && !(s_mn == 1 && mn > 1 && s_k != 1);
  const int64_t inner = k_major ? k : mn, outer = k_major ? mn : k;
  const int64_t ld = (inner + 3) / 4 * 4;
  int rc = sk_malloc((size_t)(outer * ld) * sizeof(float), (void **)out);
  if (rc) return rc;
  sk_array s, d;
  s.data = const_cast<void *>(src); s.dtype = SK_F32; s.ndim = 2;
  s.shape[0] = outer; s.shape[1] = inner;
  s.strides[0] = k_major ? s_mn : s_k; s.strides[1] = k_major ? s_k : s_mn;
  d.data = *out; d.dtype = SK_F32; d.ndim = 2;
  d.shape[0] = outer; d.shape[1] = inner; d.strides[0] = ld; d.strides[1] = 1;
  if ((rc = sk_copy(&s, &d))) return rc;
  *o_mn = k_major ? ld : 1;
  *o_k = k_major ? 1 : ld;
  return SK_OK;
}

static int run_repitched(const GemmProblem &g0) {
  GemmProblem g = g0;
  float *ta = nullptr, *tb = nullptr;
  int rc = SK_OK;
  for (int64_t bz = 0; bz < g.batch && rc == SK_OK; ++bz) {
    GemmProblem gi = g;
    gi.batch = 1;
    gi.a = (const float *)g.a + bz * g.sa_b;
    gi.b = (const float *)g.b + bz * g.sb_b;
    gi.c = g.c + bz * g.sc_b;
    if (!tc_operand_ok(g.M, g.K, g.sa_m, g.sa_k, 4, gi.a)) {
      if ((rc = repitch(gi.a, g.M, g.K, g.sa_m, g.sa_k, &ta, &gi.sa_m, &gi.sa_k))) break;
      gi.a = ta;
    }
    if (!tc_operand_ok(g.N, g.K, g.sb_n, g.sb_k, 4, gi.b)) {
      if ((rc = repitch(gi.b, g.N, g.K, g.sb_n, g.sb_k, &tb, &gi.sb_n, &gi.sb_k))) break;
      gi.b = tb;
    }
    if (tc_supported(gi, SK_MM_TF32X3)) rc = launch_gemm_tc(gi, SK_MM_TF32X3);
    else rc = run_one(gi, SK_MM_SIMT);
    if (ta) { sk_free(ta); ta = nullptr; }
    if (tb) { sk_free(tb); tb = nullptr; }
  }
  if (ta) sk_free(ta);
  if (tb) sk_free(tb);
  return rc;
}

static int run_one(const GemmProblem &g0, int algo) {
  GemmProblem g = g0;
  if (g.M == 0 || g.N == 0 || g.batch == 0) return SK_OK;
  if (g.K == 0) {
    // empty inner dimension: result is epilogue(0)
    set_error("matmul: K == 0 is not supported");
    return SK_ERR_UNSUPPORTED;
  }
  if (g.a_dtype == SK_BF16 || g.b_dtype == SK_BF16) {
    SK_REQUIRE(g.a_dtype == SK_BF16 && g.b_dtype == SK_BF16, "matmul: mixed bf16/fp32 operands");
    SK_REQUIRE(algo == SK_MM_AUTO || algo == SK_MM_BF16, "matmul: bf16 operands need SK_MM_BF16");
    if (!tc_supported(g, SK_MM_BF16)) {
      set_error("matmul(bf16): shape/strides not supported by the tcgen05 path (M=%lld N=%lld K=%lld)",
                (long long)g.M, (long long)g.N, (long long)g.K);
      return SK_ERR_UNSUPPORTED;
    }
    return launch_gemm_tc(g, SK_MM_BF16);
  }
  if (g.a_dtype == SK_F64 || g.b_dtype == SK_F64) {
    // float64 operands (np.matmul of float64 / integer Tensors, forward.pyx:172-178): CUDA-core DFMA
    SK_REQUIRE(g.a_dtype == SK_F64 && g.b_dtype == SK_F64, "matmul: mixed float64 / float32 operands");
    SK_REQUIRE(g.epilogue == SK_EPI_NONE, "matmul(float64): no fused epilogue");
    return launch_gemm_f64(g);
  }
  SK_REQUIRE(g.a_dtype == SK_F32 && g.b_dtype == SK_F32, "matmul: operands must be float32, float64 or bf16");
  if (algo == SK_MM_AUTO) {
    // SOKET_B200_FP32_GEMM = f16x3 (default) | tf32x3 selects the fp32-parity tensor-core scheme
    static const bool want_f16x3 = !(getenv("SOKET_B200_FP32_GEMM") && !strcmp(getenv("SOKET_B200_FP32_GEMM"), "tf32x3"));
    if (!tc_profitable(g)) algo = SK_MM_SIMT;
    else if (want_f16x3 && tc_supported(g, SK_MM_F16X3)) algo = SK_MM_F16X3;
    else if (tc_supported(g, SK_MM_TF32X3)) algo = SK_MM_TF32X3;
    else return run_repitched(g);   // large problem whose row pitch TMA cannot describe
  }
  if (algo == SK_MM_SIMT) {
    MMArgs p;
    p.a = (const float *)g.a; p.b = (const float *)g.b; p.c = g.c; p.bias = g.bias;
    p.M = g.M; p.N = g.N; p.K = g.K;
    p.sa_m = g.sa_m; p.sa_k = g.sa_k; p.sb_k = g.sb_k; p.sb_n = g.sb_n; p.ldc = g.ldc;
    p.batch = g.batch; p.sa_b = g.sa_b; p.sb_b = g.sb_b; p.sc_b = g.sc_b;
    p.epilogue = g.epilogue;
    return launch_gemm_simt(p);
  }
  if (!tc_supported(g, algo)) {
    set_error("matmul: tcgen05 path does not support this problem (M=%lld N=%lld K=%lld, strides a(%lld,%lld) b(%lld,%lld))",
              (long long)g.M, (long long)g.N, (long long)g.K, (long long)g.sa_m, (long long)g.sa_k,
              (long long)g.sb_k, (long long)g.sb_n);
    return SK_ERR_UNSUPPORTED;
  }
  return launch_gemm_tc(g, algo);
}

static int matmul_impl(const sk_array *a, const sk_array *b, const sk_array *bias, sk_array *out,
                       int epilogue, int algo) {
  int rc;
  if ((rc = ensure_init())) return rc;
  GemmProblem g;
  memset(&g, 0, sizeof(g));
  if ((rc = parse_2d(a, b, out, g))) return rc;
  g.a = a->data; g.b = b->data; g.c = (float *)out->data;
  g.a_dtype = a->dtype; g.b_dtype = b->dtype;
  g.epilogue = epilogue;
  if (epilogue == SK_EPI_BIAS || epilogue == SK_EPI_BIAS_RELU) {
    SK_REQUIRE(bias && bias->dtype == SK_F32 && numel(bias) == g.N &&
                   (g.N == 1 || bias->strides[bias->ndim - 1] == 1),
               "linear: bias must be a contiguous float32 vector of length N");
    g.bias = (const float *)bias->data;
  }
  // leading (batch) dims: broadcast operands against out's leading dims
  const int nb = out->ndim - 2;
  int64_t bshape[SK_MAX_NDIM], sa[SK_MAX_NDIM], sb[SK_MAX_NDIM], sc[SK_MAX_NDIM];
  for (int i = 0; i < nb; ++i) {
    bshape[i] = out->shape[i];
    sc[i] = out->strides[i];
    int ia = i - (nb - (a->ndim - 2)), ib = i - (nb - (b->ndim - 2));
    sa[i] = (ia >= 0 && a->shape[ia] != 1) ? a->strides[ia] : 0;
    sb[i] = (ib >= 0 && b->shape[ib] != 1) ? b->strides[ib] : 0;
    if (ia >= 0) SK_REQUIRE(a->shape[ia] == 1 || a->shape[ia] == bshape[i], "matmul: batch dims do not broadcast");
    if (ib >= 0) SK_REQUIRE(b->shape[ib] == 1 || b->shape[ib] == bshape[i], "matmul: batch dims do not broadcast");
  }
  const int64_t *strs[3] = {sa, sb, sc};
  Collapsed<3> c;
  collapse_dims<3>(nb, bshape, strs, c);
  // innermost collapsed dim becomes the kernel's batch dim; outer ones loop on the host
  const int last = c.ndim - 1;
  g.batch = c.shape[last];
  g.sa_b = c.strides[0][last]; g.sb_b = c.strides[1][last]; g.sc_b = c.strides[2][last];
  int64_t outer = 1;
  for (int i = 0; i < last; ++i) outer *= c.shape[i];
  const int esz_a = dtype_size(a->dtype), esz_b = dtype_size(b->dtype);
  for (int64_t o = 0; o < outer; ++o) {
    int64_t rem = o, oa = 0, ob = 0, oc = 0;
    for (int k = last - 1; k >= 0; --k) {
      int64_t idx = rem % c.shape[k];
      rem /= c.shape[k];
      oa += idx * c.strides[0][k]; ob += idx * c.strides[1][k]; oc += idx * c.strides[2][k];
    }
    GemmProblem gi = g;
    gi.a = (const char *)g.a + oa * esz_a;
    gi.b = (const char *)g.b + ob * esz_b;
    gi.c = (float *)((char *)g.c + oc * dtype_size(out->dtype));
    if ((rc = run_one(gi, algo))) return rc;
  }
  return SK_OK;
}

}  // namespace sk

using namespace sk;

extern "C" {

int sk_matmul(const sk_array *a, const sk_array *b, sk_array *out, int algo) {
  return matmul_impl(a, b, nullptr, out, SK_EPI_NONE, algo);
}

int sk_linear_fwd(const sk_array *x, const sk_array *w, const sk_array *bias, sk_array *out,
                  int epilogue, int algo) {
  SK_REQUIRE(epilogue >= SK_EPI_NONE && epilogue <= SK_EPI_RELU, "linear: bad epilogue %d", epilogue);
  return matmul_impl(x, w, bias, out, epilogue, algo);
}

int sk_split_f16(const float *x, int64_t rows, int64_t cols, int64_t ldx, const unsigned int *amax_bits, void *hi,
                 void *lo, int64_t ldh, float *scale4, float *colsum_out) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(x && hi && lo && scale4, "sk_split_f16: null pointer");
  SK_REQUIRE(rows > 0 && cols > 0, "sk_split_f16: empty matrix");
  SK_REQUIRE(ldx >= cols && ldh >= cols && ldh % 8 == 0 && ldh == (cols + 7) / 8 * 8,
             "sk_split_f16: ldh must be cols rounded up to a multiple of 8 (got %lld for %lld columns)",
             (long long)ldh, (long long)cols);
  SK_REQUIRE(ldx % 4 == 0 || rows == 1, "sk_split_f16: source pitch must be a multiple of 4 elements");
  SK_REQUIRE(((((uintptr_t)x) | ((uintptr_t)hi) | ((uintptr_t)lo) | ((uintptr_t)scale4)) & 15) == 0,
             "sk_split_f16: pointers must be 16-byte aligned");
  ProfScope pp(SK_PROF_GEMM_PREP, (double)rows * (double)cols * (amax_bits ? 8.0 : 12.0));
  return split_f16_tensor(x, ldx, rows, cols, (const uint32_t *)amax_bits, (__half *)hi, (__half *)lo, ldh, scale4,
                          colsum_out);
}

int sk_gemm_f16x3_supported(int64_t M, int64_t N, int64_t K) { return gemm_f16x3_shape_ok(M, N, K) ? 1 : 0; }

int sk_gemm_f16x3(const sk_split_operand *a, const sk_split_operand *b, float *c, int64_t ldc, int64_t M, int64_t N,
                  int64_t K, const float *bias, int epilogue, int accumulate) {
  int rc;
  if ((rc = ensure_init())) return rc;
  return gemm_f16x3_presplit(a, b, c, ldc, M, N, K, bias, epilogue, accumulate);
}

static int linear_bwd_impl(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw,
                           float *db);

int sk_linear_bwd(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw) {
  return linear_bwd_impl(adj, x, w, dx, dw, nullptr);
}

int sk_linear_bwd_bias(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw,
                       float *db) {
  SK_REQUIRE(db != nullptr, "sk_linear_bwd_bias: null db");
  return linear_bwd_impl(adj, x, w, dx, dw, db);
}

static int linear_bwd_impl(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw,
                           float *db) {
  int rc;
  if ((rc = ensure_init())) return rc;
  SK_REQUIRE(adj && x && w && dx && dw, "sk_linear_bwd: null array");
  SK_REQUIRE(adj->ndim == 2 && x->ndim == 2 && w->ndim == 2 && dx->ndim == 2 && dw->ndim == 2,
             "sk_linear_bwd: all operands must be 2-D");
  const int64_t Bn = adj->shape[0], O = adj->shape[1], I = x->shape[1];
  SK_REQUIRE(x->shape[0] == Bn && w->shape[0] == I && w->shape[1] == O && dx->shape[0] == Bn && dx->shape[1] == I &&
                 dw->shape[0] == I && dw->shape[1] == O,
             "sk_linear_bwd: shapes must be adj (B,O), x (B,I), w (I,O), dx (B,I), dw (I,O)");
  static const bool want_f16x3 = !(getenv("SOKET_B200_FP32_GEMM") && !strcmp(getenv("SOKET_B200_FP32_GEMM"), "tf32x3"));
  const sk_array *all[5] = {adj, x, w, dx, dw};
  bool plain = want_f16x3;
  for (const sk_array *a : all) plain = plain && a->dtype == SK_F32 && is_contiguous(a);
  if (plain) {
    bool done = false;
    rc = linear_bwd_f16x3((const float *)adj->data, (const float *)x->data, (const float *)w->data,
                          (float *)dx->data, (float *)dw->data, db, Bn, I, O, &done);
    if (rc || done) return rc;
  }
  if (db) {   // general path: the bias gradient is its own column-sum pass
    SK_REQUIRE(adj->dtype == SK_F32 && is_contiguous(adj), "sk_linear_bwd_bias: adj must be contiguous float32");
    if ((rc = sk_colsum((const float *)adj->data, nullptr, db, Bn, O))) return rc;
  }
  // general path: two GEMMs on .T views (backward.pyx:720-736)
  sk_array wt = *w, xt = *x;
  wt.shape[0] = w->shape[1]; wt.shape[1] = w->shape[0]; wt.strides[0] = w->strides[1]; wt.strides[1] = w->strides[0];
  xt.shape[0] = x->shape[1]; xt.shape[1] = x->shape[0]; xt.strides[0] = x->strides[1]; xt.strides[1] = x->strides[0];
  if ((rc = matmul_impl(adj, &wt, nullptr, dx, SK_EPI_NONE, SK_MM_AUTO))) return rc;
  return matmul_impl(&xt, adj, nullptr, dw, SK_EPI_NONE, SK_MM_AUTO);
}

}  // extern "C"
