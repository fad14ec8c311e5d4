/*
 * soket_b200.h -- C ABI of libsoketb200.so, the B200 (sm_100a) device backend
 * that sits where CuPy sits in singul4ri7y/soket.
 *
 * Every entry point takes plain pointers / sizes / small POD structs, returns
 * an int status (SK_OK == 0) and never throws.  On failure the message is
 * available from sk_last_error().  All launches go to ONE per-process stream
 * (sk_stream()), are asynchronous, and surface execution errors at the next
 * sk_sync()/sk_d2h().  The library never frees caller memory and never keeps
 * a caller pointer past the call.  Single-threaded use (the reference holds the
 * GIL and is not thread-safe: soket/backend/device.pyx:261, tensor.pyx:24).
 *
 * "replaces:" comments cite the reference interface (file:line under the
 * soket tree) that each entry point stands in for.
 */
#ifndef SOKET_B200_H
#define SOKET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_OK 0
#define SK_ERR_CUDA 1
#define SK_ERR_ARG 2
#define SK_ERR_UNSUPPORTED 3
#define SK_ERR_NCCL 4
#define SK_ERR_OOM 5
#define SK_ERR_INDEX 6 /* an index read by a kernel was out of range (reported at the next sync point) */

#define SK_MAX_NDIM 8

/* dtype tags: same 12 names as soket/dtype.pyx:12-14, plus bf16 (GEMM input only). */
typedef enum {
  SK_BOOL = 0,
  SK_I8 = 1,
  SK_U8 = 2,
  SK_I16 = 3,
  SK_U16 = 4,
  SK_I32 = 5,
  SK_U32 = 6,
  SK_I64 = 7,
  SK_U64 = 8,
  SK_F16 = 9,
  SK_F32 = 10,
  SK_F64 = 11,
  SK_BF16 = 12
} sk_dtype;

/* Strided view of device memory.  `data` points at element [0,...,0];
 * strides are in ELEMENTS (may be 0 = broadcast, or negative). */
typedef struct {
  void *data;
  int32_t dtype;
  int32_t ndim;
  int64_t shape[SK_MAX_NDIM];
  int64_t strides[SK_MAX_NDIM];
} sk_array;

/* ------------------------------------------------------------------ runtime
 * replaces: cupy.cuda.Device(id).use()/synchronize() as called from
 * soket/backend/device.pyx:56-58,188-198; cupy's memory pool; cupy.asnumpy /
 * cupy.array H2D-D2H (soket/tensor/tensor.pyx:384-442). */
int sk_init(int device);
int sk_device_count(int *count);
int sk_current_device(int *device);
int sk_sync(void);
const char *sk_last_error(void);
const char *sk_version(void);
void *sk_stream(void); /* cudaStream_t of the compute stream */

int sk_malloc(size_t nbytes, void **ptr); /* stream-ordered caching allocator */
int sk_free(void *ptr);
int sk_empty_cache(void);
int sk_mem_stats(size_t *in_use, size_t *reserved, size_t *peak_in_use);
int sk_host_alloc(size_t nbytes, void **ptr); /* pinned host memory */
int sk_host_free(void *ptr);
int sk_h2d(void *dst, const void *src, size_t nbytes);       /* sync on return */
int sk_d2h(void *dst, const void *src, size_t nbytes);       /* sync on return */
int sk_h2d_async(void *dst, const void *src, size_t nbytes); /* src must be pinned */
int sk_d2h_async(void *dst, const void *src, size_t nbytes); /* dst must be pinned */
int sk_d2d(void *dst, const void *src, size_t nbytes);
int sk_memset(void *dst, int byte, size_t nbytes);

/* timing + launch accounting (bench.py: CUDA events on the launching stream) */
/* host -> device input prefetch on a dedicated copy stream: ordered after the compute work
 * queued so far, overlapping the work queued next; sk_prefetch_wait() joins it back into the
 * compute stream (Tensor(host_array) of the reference, tensor.pyx:386-442, made asynchronous) */
int sk_h2d_prefetch(void *dst, const void *src, size_t nbytes);
int sk_prefetch_wait(void);
int sk_event_create(void **ev);
int sk_event_record(void *ev);
int sk_event_sync(void *ev);
int sk_event_elapsed_ms(void *start, void *stop, float *ms);
int sk_event_destroy(void *ev);
/* The process's streams.  Everything launches on the compute stream unless sk_launch_stream() selects
 * another one (a host-side switch, single-threaded like the rest of the API): data-parallel training
 * issues its bucketed all-reduces on the comm stream and the per-bucket optimizer update on the
 * optimizer stream while backward keeps the compute stream busy. */
typedef enum { SK_STREAM_COMPUTE = 0, SK_STREAM_COMM = 1, SK_STREAM_COPY = 2, SK_STREAM_OPT = 3 } sk_stream_id;
int sk_launch_stream(int stream_id);                     /* later launches go to this stream */
int sk_event_record_on(void *ev, int stream_id);
int sk_stream_wait_event(int stream_id, void *ev);       /* stream waits for the event (device side) */
uint64_t sk_launch_count(void); /* kernels this library launched so far */
int sk_flush_l2(void);          /* overwrite a >L2 scratch buffer */

/* Per-kernel-family timing for the roofline numbers: when enabled, every launch of
 * an instrumented family is bracketed by CUDA events on the compute stream;
 * sk_prof_collect() synchronises and returns, for one family, the launch count,
 * the summed device time and the summed ALGORITHMIC work (flops for GEMMs, bytes
 * for memory-bound kernels) since the last sk_prof_reset(). */
typedef enum {
  SK_PROF_GEMM_TC = 0,
  SK_PROF_GEMM_SIMT = 1,
  SK_PROF_LN_FWD = 2,
  SK_PROF_LN_BWD = 3,
  SK_PROF_EWISE = 4,
  SK_PROF_REDUCE = 5,
  SK_PROF_OPTIM = 6,
  SK_PROF_COPY = 7,
  SK_PROF_BN = 8,
  SK_PROF_LOSS = 9,
  SK_PROF_GEMM_PREP = 10, /* operand splits (fp16 hi/lo, tf32 lo) feeding the tcgen05 GEMMs; work = 8 B/elem */
  SK_PROF_NUM = 11
} sk_prof_family;
int sk_prof_enable(int on);
int sk_prof_reset(void);
int sk_prof_collect(int family, int64_t *launches, double *total_ms, double *total_work);

/* CUDA-graph capture of a static-shape step.
 * Allocation arenas: between sk_arena_begin(id) and sk_arena_end() every sk_malloc is served
 * from (and every block returns to) arena `id`, so the buffers a graph was captured with stay
 * reserved for its replays.  During a capture a miss in the arena is an error (no cudaMalloc
 * inside a capture): run the step once or twice inside the arena first.
 * sk_rng_epoch_advance(): bumps the device-side counter the dropout kernels mix into their
 * seeds at run time -- captured as the first node, it gives every replay fresh masks. */
int sk_arena_create(int *arena);
int sk_arena_begin(int arena);
int sk_arena_end(void);
int sk_arena_destroy(int arena);
int sk_rng_epoch_advance(void);
int sk_graph_begin(void);
int sk_graph_capturing(void);   /* 1 between sk_graph_begin and sk_graph_end */
int sk_graph_end(void **graph_exec);
int sk_graph_launch(void *graph_exec);
int sk_graph_destroy(void *graph_exec);

/* ------------------------------------------------------- elementwise family
 * replaces: intern-table slots _ADD.._POW, _NEG, _MAXIMUM, _EXP, _LOG,
 * _EQUAL.._LESS_EQUAL, _ARRAY/_COPY, _BCASTTO/_TRANSPOSE/_RESHAPE when they
 * must be materialised (soket/tensor/ops/intern.pyx:45-76; call shapes in
 * soket/tensor/ops/forward.pyx:7-221). NumPy broadcasting is expressed by the
 * caller through zero strides; `out` is written in its own dtype. */
typedef enum {
  SK_OP_ADD = 0,
  SK_OP_SUB = 1,
  SK_OP_MUL = 2,
  SK_OP_DIV = 3,
  SK_OP_POW = 4,
  SK_OP_MAXIMUM = 5,
  SK_OP_MINIMUM = 6,
  /* comparisons: out dtype bool */
  SK_OP_EQ = 16,
  SK_OP_NE = 17,
  SK_OP_GT = 18,
  SK_OP_GE = 19,
  SK_OP_LT = 20,
  SK_OP_LE = 21
} sk_binary_op;

typedef enum {
  SK_UOP_NEG = 0,
  SK_UOP_EXP = 1,
  SK_UOP_LOG = 2,
  SK_UOP_SQRT = 3,
  SK_UOP_RELU = 4, /* maximum(x, 0) fast path: forward.pyx:206-209 */
  SK_UOP_ABS = 5
} sk_unary_op;

/* out = a (op) b ; a, b, out share shape (broadcast dims have stride 0). */
int sk_ewise_binary(int op, const sk_array *a, const sk_array *b, sk_array *out);
/* out = a (op) s, or s (op) a when reverse != 0.  The scalar is a weak Python
 * scalar (NEP 50): it adopts the array's kind, so it is passed as a double and
 * an int64 (scalar_is_int selects). */
int sk_ewise_scalar(int op, const sk_array *a, double fscalar, int64_t iscalar,
                    int scalar_is_int, int reverse, sk_array *out);
int sk_ewise_unary(int op, const sk_array *a, sk_array *out);
/* dst[...] = cast(src[...]) for arbitrary strides on both sides: compaction,
 * astype, broadcast materialisation, slice __setitem__ (tensor.pyx:948). */
int sk_copy(const sk_array *src, sk_array *dst);
int sk_fill(sk_array *dst, double fvalue, int64_t ivalue, int value_is_int);
/* A CHAIN of float32 elementwise / scalar / unary ops as ONE launch (lazy mode, tensor.pyx:24-51,
 * 790-810: with soket.lazy() the graph is known before anything runs, so a run of elementwise nodes
 * is evaluated in registers instead of one HBM pass per node).  The chain is a small accumulator
 * program interpreted per element -- every step is the same single fp32 operation the unfused kernel
 * would have performed, in the same order, so the result is bit-identical to the op-by-op evaluation:
 *   SK_F_LOAD   acc = operand               SK_F_STORE  temp[idx] = acc
 *   SK_F_BIN    acc = acc (sub) operand     (rev: operand (sub) acc), sub = sk_binary_op ADD..MINIMUM
 *   SK_F_UN     acc = (sub) acc             sub = sk_unary_op
 * operand = an input array (src SK_F_IN: full-size, a last-axis vector of `cols` elements, or a single
 * element), a constant (SK_F_CONST) or a temporary (SK_F_TEMP, 3 of them). */
#define SK_FUSED_MAX_OPS 48
#define SK_FUSED_MAX_INPUTS 8
typedef enum { SK_F_LOAD = 0, SK_F_STORE = 1, SK_F_BIN = 2, SK_F_UN = 3 } sk_fused_code;
typedef enum { SK_F_IN = 1, SK_F_CONST = 2, SK_F_TEMP = 3 } sk_fused_src;
typedef enum { SK_F_FULL = 0, SK_F_VECTOR = 1, SK_F_SINGLE = 2 } sk_fused_input_kind;
typedef struct {
  int n_ops, n_in;
  unsigned char code[SK_FUSED_MAX_OPS], sub[SK_FUSED_MAX_OPS], src[SK_FUSED_MAX_OPS], idx[SK_FUSED_MAX_OPS],
      rev[SK_FUSED_MAX_OPS];
  float cst[SK_FUSED_MAX_OPS];          /* the constant of a SK_F_CONST operand */
  const float *in[SK_FUSED_MAX_INPUTS];
  int in_kind[SK_FUSED_MAX_INPUTS];
  int64_t n, cols;                      /* elements of the result; length of the last axis */
} sk_fused_program;
int sk_ewise_fused(const sk_fused_program *prog, float *out);
/* relu backward: out = (x > 0) * adj  (backward.pyx:849-874) in one pass */
int sk_relu_bwd(const sk_array *x, const sk_array *adj, sk_array *out);

/* ---------------------------------------------------------------- reductions
 * replaces: _SUM/_MEAN/_MAX/_MIN/_ARGMAX/_ARGMIN (intern.pyx:52-57), call
 * convention f(x, axes, dtype, out, keepdims) (forward.pyx:128-170).
 * `axes_mask` bit i set => axis i of `in` is reduced.  `out` is the contiguous
 * result WITHOUT the reduced axes (keepdims is a host-side reshape). */
typedef enum { SK_RED_SUM = 0, SK_RED_MEAN = 1, SK_RED_MAX = 2, SK_RED_MIN = 3 } sk_reduce_op;
int sk_reduce(int op, const sk_array *in, uint32_t axes_mask, sk_array *out);
/* axis < 0: over the flattened array.  out dtype int64 (NumPy) or int32. */
int sk_argreduce(int is_min, const sk_array *in, int axis, sk_array *out);

/* ------------------------------------------------------------------ indexing
 * replaces: ndarray.__getitem__ with an integer array (device.pyx:236-239,
 * `eye(C)[labels]`), eye (device.pyx:69), np.stack (util.pyx:36). */
int sk_gather_rows(const sk_array *src, const sk_array *index, sk_array *out);
int sk_eye(sk_array *out, int64_t k);
int sk_one_hot(const sk_array *labels, sk_array *out); /* out[b, labels[b]] = 1 else 0 */

/* ----------------------------------------------------------------------- RNG
 * replaces: cupy.random.uniform/normal/binomial(1,p) (device.pyx:64-66,
 * 204-224).  Philox4x32-10, counter-based; stream = (seed, offset). */
int sk_rng_seed(uint64_t seed);
int sk_rng_uniform(sk_array *out, double low, double high);
int sk_rng_normal(sk_array *out, double mean, double std);
int sk_rng_bernoulli(sk_array *out, double p);

/* -------------------------------------------------------------------- matmul
 * replaces: _MATMUL (intern.pyx:63): np.matmul(x, y, dtype='float32') with
 * row-major or transposed (.T view) 2-D operands (forward.pyx:172-178,
 * backward.pyx:704-742) and leading batch dims.
 * a: (..., M, K)   b: (..., K, N)   out: (..., M, N) contiguous fp32.
 * Operand strides select K-major / MN-major in the tensor-core path; anything
 * else is compacted by the caller first.
 *   SK_MM_AUTO      tcgen05 fp16x3 (CTA-pair tiles) when shapes allow, else 3xTF32, else SIMT fp32
 *   SK_MM_F16X3     tcgen05 kind::f16 on fp16 hi/lo splits of the fp32 operands with one
 *                   power-of-two scale per row of a / column of b (fp32 parity, 2x 3xTF32)
 *   SK_MM_SIMT      CUDA-core fp32 FFMA (exact fp32 products; validation path)
 *   SK_MM_TF32X3    tcgen05 kind::tf32, error-compensated hi/lo split
 *   SK_MM_TF32      tcgen05 kind::tf32 single pass (1e-3 class; not parity)
 *   SK_MM_BF16      tcgen05 kind::f16 on bf16 operands (a,b dtype SK_BF16) */
typedef enum { SK_MM_AUTO = 0, SK_MM_SIMT = 1, SK_MM_TF32X3 = 2, SK_MM_TF32 = 3, SK_MM_BF16 = 4, SK_MM_F16X3 = 5 } sk_mm_algo;
typedef enum {
  SK_EPI_NONE = 0,
  SK_EPI_BIAS = 1,      /* + bias[n]                (prototypes.pyx:108-115) */
  SK_EPI_BIAS_RELU = 2, /* relu(. + bias[n])        (+ prototypes.pyx:302)   */
  SK_EPI_RELU = 3
} sk_mm_epilogue;
int sk_matmul(const sk_array *a, const sk_array *b, sk_array *out, int algo);
/* fused Linear(+ReLU): out = epi(x @ w + bias). bias may be NULL for EPI_NONE/RELU */
int sk_linear_fwd(const sk_array *x, const sk_array *w, const sk_array *bias,
                  sk_array *out, int epilogue, int algo);
/* Pre-split operands of the fp16x3 GEMM.  The fp32-parity matmul (`_MATMUL`, forward.pyx:172-178) runs on
 * tcgen05 as three fp16 MMAs over X * scale = hi + lo; sk_matmul / sk_linear_* split their operands inside
 * every call.  A training step uses each matrix in several GEMMs (a layer's input in forward and in
 * dW = X.T @ adj, backward.pyx:734; its weight in forward and in dX = adj @ W.T, backward.pyx:722; the
 * adjoint in dX and dW), so the resident path splits each ONCE, with one power-of-two scale for the
 * whole matrix -- which is what lets the same hi / lo pair be consumed K-major by one GEMM and
 * MN-major by another -- or has the producing kernel emit the split (sk_layernorm_*_ex).
 *   scale4: device float[4] = {scale, 1/scale, |max| (or the bound the scale came from), 0}. */
typedef struct {
  const void *hi, *lo;   /* fp16, identical layout: stored rows of `ld` elements (ld % 8 == 0) */
  int64_t ld;
  int mn_major;          /* 0: a stored row is one M (A) / N (B) index with K contiguous; 1: a stored row is one K */
  const float *scale;    /* the matrix's scale4 */
} sk_split_operand;
/* x (rows, cols) pitch ldx -> hi, lo (pitch ldh = cols rounded up to 8, padding zero) + scale4.  amax_bits:
 * device word with the bit pattern of max |x| or of an upper bound (NULL: computed by an extra pass).
 * colsum_out (or NULL): the column sums of x from the same pass (the bias gradient, autodiff.pyx:84). */
int sk_split_f16(const float *x, int64_t rows, int64_t cols, int64_t ldx, const unsigned int *amax_bits, void *hi,
                 void *lo, int64_t ldh, float *scale4, float *colsum_out);
/* c (M, N) pitch ldc = epi(A @ B [+ c if accumulate]) with A = M x K, B = K x N given as split operands.
 * Needs M >= 256, N >= 128, K >= 64 (sk_gemm_f16x3_supported).  epilogue: sk_mm_epilogue. */
int sk_gemm_f16x3_supported(int64_t M, int64_t N, int64_t K);
int sk_gemm_f16x3(const sk_split_operand *a, const sk_split_operand *b, float *c, int64_t ldc, int64_t M, int64_t N,
                  int64_t K, const float *bias, int epilogue, int accumulate);
/* Linear backward (backward.pyx:704-742 for y = x @ w): dx (B,I) = adj (B,O) @ w(I,O).T and
 * dw (I,O) = x(B,I).T @ adj in one call, so that the fp16x3 path splits adj once for both
 * GEMMs; falls back to two sk_matmul calls on .T views for other shapes / layouts. */
int sk_linear_bwd(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw);
/* the same plus the bias gradient db[o] = sum_b adj[b, o] (autodiff.pyx:84-90; contiguous float32
 * vector of O elements): on the fp16x3 path it is a by-product of the pass that splits adj. */
int sk_linear_bwd_bias(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx,
                       sk_array *dw, float *db);
/* fp32 -> bf16 (RNE) cast for the bf16 sweep */
int sk_cast_bf16(const sk_array *src, sk_array *dst);

/* ------------------------------------------------------------ fused nn kernels
 * replaces the op SEQUENCES of forward.pyx:224-353 / backward.pyx:959-1132. */
/* Optional by-products of the LayerNorm kernels for the fp16x3 GEMM that consumes their result (the
 * Linear that follows a LayerNorm(+ReLU)(+Dropout) or residual LayerNorm in model.py:24-37; the Linear
 * backward that consumes a LayerNorm backward's dx).  Used by the *_ex entry points; NULL = none.
 *   forward : split_hi / split_lo (fp16, rows x cols, cols % 8 == 0) + split_scale (device float[4])
 *             receive the OUTPUT as X * scale = hi + lo.  The scale is chosen before any element exists,
 *             from |gamma * norm + beta| <= max|gamma| sqrt(cols) + max|beta| (+ the residual's bound
 *             residual_scale[2], x 1/keep under dropout) -- no extra pass, no second kernel.
 *   backward: dx_absmax (device word, zeroed by the caller) receives the bit pattern of max |dx|. */
typedef struct {
  void *split_hi, *split_lo;
  float *split_scale;
  const float *residual_scale;
  unsigned int *dx_absmax;
} sk_ln_extras;
/* LayerNorm over the last axis of a contiguous (rows, cols) fp32 matrix.
 * y = gamma * ((x-mean) * rstd) + beta ; biased variance ; saves mean/rstd.
 * relu != 0 fuses the following ReLU; residual != NULL fuses
 * relu(residual + LN(x)) (model.py:34-37, prototypes.pyx:272-273). */
int sk_layernorm_fwd(const float *x, const float *gamma, const float *beta,
                     const float *residual, float *y, float *mean, float *rstd,
                     int64_t rows, int64_t cols, float eps, int relu);
int sk_layernorm_fwd_ex(const float *x, const float *gamma, const float *beta, const float *residual,
                        float *y, float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                        int relu, const sk_ln_extras *extras);
/* dX, dgamma, dbeta (per-group partials column-reduced internally).
 * mask_mode: 0 none; 1 the LN output fed a ReLU directly: the mask (LN(x) > 0)
 * is recomputed from x/mean/rstd/gamma/beta, nothing extra is read; 2 the mask
 * is (y_out > 0) for the fused relu(residual + LN(x)) output.  dresidual (may be
 * NULL) receives the masked adjoint = gradient of the residual branch.
 * Quirk kept: dgamma/dbeta are summed over axis 0 (backward.pyx:1052-1055). */
int sk_layernorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                     const float *mean, const float *rstd, const float *y_out, int mask_mode,
                     float *dx, float *dgamma, float *dbeta, float *dresidual,
                     int64_t rows, int64_t cols);
int sk_layernorm_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                        const float *mean, const float *rstd, const float *y_out, int mask_mode,
                        float *dx, float *dgamma, float *dbeta, float *dresidual, int64_t rows,
                        int64_t cols, const sk_ln_extras *extras);
/* LayerNorm (+ReLU) followed by Dropout in ONE pass each way (the Linear - LayerNorm - ReLU -
 * Dropout run of the residual block, model.py:24-31; prototypes.pyx:746-760 for the dropout):
 * y = (relu(LN(x)) * mask) * (1/keep), mask ~ Bernoulli(keep) identified by *seed (the same draw
 * sk_dropout_fwd_seeded would make); the backward takes the adjoint of y, regenerates the mask
 * from seed ((adj * r_keep) * mask), recomputes the ReLU mask and runs the LayerNorm backward.
 * Bit-identical to sk_layernorm_fwd + sk_dropout_fwd_seeded / sk_dropout_bwd + sk_layernorm_bwd. */
int sk_layernorm_dropout_fwd(const float *x, const float *gamma, const float *beta, float *y,
                             float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                             int relu, float keep, uint64_t *seed);
int sk_layernorm_dropout_fwd_ex(const float *x, const float *gamma, const float *beta, float *y, float *mean,
                                float *rstd, int64_t rows, int64_t cols, float eps, int relu, float keep,
                                uint64_t *seed_out, const sk_ln_extras *extras);
int sk_layernorm_dropout_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                             const float *mean, const float *rstd, int relu, float keep, float r_keep,
                             uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                             int64_t cols);
int sk_layernorm_dropout_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                                const float *mean, const float *rstd, int relu, float keep, float r_keep,
                                uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                                int64_t cols, const sk_ln_extras *extras);
/* BatchNorm1d over axis 0 of a contiguous (rows, cols) fp32 matrix, training
 * mode (always: quirk Q4, forward.pyx:281), biased variance, running stats
 * rm = (1-m) rm + m mean (forward.pyx:308-318). */
int sk_batchnorm_fwd(const float *x, const float *gamma, const float *beta,
                     float *y, float *mean, float *rstd, float *running_mean,
                     float *running_var, int64_t rows, int64_t cols, float eps,
                     float momentum, int relu);
int sk_batchnorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                     const float *mean, const float *rstd, const float *y_out, int mask_mode,
                     float *dx, float *dgamma, float *dbeta, int64_t rows, int64_t cols);
/* softmax cross-entropy, mean reduction, integer labels (no one-hot):
 * loss = mean_b(logsumexp(x_b) - x_b[y_b]) ; dx = (softmax(x) - onehot) / B.
 * labels dtype u8/i32/i64.  (forward.pyx:250-271, backward.pyx:959-1022) */
int sk_softmax_ce_fwd_bwd(const float *logits, const void *labels, int label_dtype,
                          float *loss, float *dlogits, float *row_loss,
                          int64_t rows, int64_t classes);
/* out = relu(a + b)  (Residual + outer ReLU) */
int sk_add_relu(const float *a, const float *b, float *out, int64_t n);
/* dropout: out = x * mask * (1/keep), mask ~ Bernoulli(keep) written as fp32 */
int sk_dropout_fwd(const float *x, float *out, float *mask, int64_t n, float keep);
/* the same without a stored mask: *seed identifies the Bernoulli draw; sk_dropout_bwd regenerates
 * the mask from it: out = (adj * r_keep) * mask  (prototypes.pyx:746-760 backward) */
int sk_dropout_fwd_seeded(const float *x, float *out, int64_t n, float keep, uint64_t *seed);
int sk_dropout_bwd(const float *adj, float *out, int64_t n, float keep, float r_keep, uint64_t seed);
/* bias gradient: out[c] = sum_r adj[r, c]  (autodiff.pyx:43-101) ;
 * y_out != NULL applies the relu mask first. */
int sk_colsum(const float *adj, const float *y_out, float *out, int64_t rows, int64_t cols);
/* grad accumulation in place: acc += part (autodiff.pyx:30-41) */
int sk_accumulate(float *acc, const float *part, int64_t n);

/* ------------------------------------------------------------------ optimizers
 * replaces: SGD.step (optim.pyx:82-131) and Adam.step (optim.pyx:201-269):
 * ONE launch over all parameter tensors (HOST arrays of device pointers; up to
 * 48 tensors per launch).  Hyper-parameters arrive as the Python doubles the
 * reference holds and are rounded to float32 exactly where NumPy would (NEP 50).  grad_scale folds the 1/W
 * of data-parallel averaging.  Arithmetic order follows the reference
 * including quirks Q2/Q3. */
int sk_sgd_step(int n_tensors, float *const *params, const float *const *grads,
                const int64_t *sizes, double lr, double weight_decay, double grad_scale);
int sk_adam_step(int n_tensors, float *const *params, const float *const *grads,
                 float *const *m, float *const *v, const int64_t *sizes, double lr,
                 double beta1, double beta2, double eps, double weight_decay,
                 double one_minus_beta1_t, double one_minus_beta2_t, int first_step,
                 double grad_scale);
/* The same update with the running products {beta1^t, beta2^t} (optim.pyx:191-195,266-269)
 * held as two doubles in DEVICE memory: the kernel forms 1 - beta^t itself (same double
 * subtraction, same float32 rounding as the host path), sk_adam_bias_advance multiplies them
 * by the betas after a step.  Nothing step-dependent is left in the kernel arguments, so an
 * Adam step can be captured in a CUDA graph and replayed. */
int sk_adam_step_dev(int n_tensors, float *const *params, const float *const *grads,
                     float *const *m, float *const *v, const int64_t *sizes, double lr,
                     double beta1, double beta2, double eps, double weight_decay,
                     const double *bias_state, int first_step, double grad_scale);
/* sk_adam_step / sk_adam_step_dev (bias_state != NULL) that also leaves the bit pattern of max |p_new| of
 * tensor i in the device word amax[i] (NULL entries: skipped; the caller zeroes the words): the scale
 * of the weight's fp16x3 operand split (sk_split_f16 amax_bits) without a pass over the weight. */
int sk_adam_step_amax(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                      float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                      double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                      const double *bias_state, int first_step, double grad_scale, unsigned int *const *amax);
/* sk_adam_step_amax with PERSISTENT |max| words and, optionally, the new weights' fp16x3 operand split
 * written by the same kernel (replaces the weight's own sk_split_f16 pass after every step:
 * 4 B/element of extra writes instead of an 8 B/element pass).
 *   amax2   device uint32[2] per tensor (NULL: tensor not tracked): [0] = bit pattern of max |p| BEFORE
 *           the update (0 = unknown: only allowed without hi), [1] = 0.  After the launch [0] holds
 *           max |p_new| and [1] is 0 again, so the same words serve the next step (and a CUDA-graph replay).
 *   hi, lo  fp16 arrays of `size` elements or NULL; scale4 = float[4] {scale, 1/scale, bound, 0}:
 *           p_new * scale = hi + lo, scale = the power of two sk_split_f16 would pick for
 *           bound = max |p_old| + update_bound -- chosen before any element is updated.
 *   update_bound  a bound of |p_new - p_old| that holds for EVERY element whatever the gradients:
 *           lr * max_t |m_hat / sqrt(v_hat)| <= lr * (1-b1)/sqrt(1-b2) * sqrt(sum_k (b1^2/b2)^k) *
 *           sqrt(1-b2^t)/(1-b1^t) (Cauchy-Schwarz on the two moving averages; optim.pyx:224-263). */
typedef struct {
    unsigned int *amax2;
    void *hi, *lo;
    float *scale4;
} sk_adam_split;
int sk_adam_step_split(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                       float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                       double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                       const double *bias_state, int first_step, double grad_scale, const sk_adam_split *splits,
                       double update_bound);
int sk_adam_bias_advance(double *bias_state, double beta1, double beta2);

/* ------------------------------------------------------------- data parallel
 * new (the reference has no collective): NCCL all-reduce(sum) of fp32
 * gradient buckets on a dedicated comm stream. */
#define SK_NCCL_ID_BYTES 128
int sk_nccl_available(void);
int sk_nccl_unique_id(char id[SK_NCCL_ID_BYTES]);
int sk_nccl_init(int rank, int world, const char id[SK_NCCL_ID_BYTES]);
int sk_nccl_allreduce(float *buf, size_t count, int on_comm_stream);
int sk_nccl_broadcast(float *buf, size_t count, int root);
int sk_nccl_wait(void); /* compute stream waits for the comm stream */
/* all-reduce(sum) in place on the given stream; ordering is the caller's (events above) */
int sk_nccl_allreduce_on(float *buf, size_t count, int stream_id);
/* any NCCL error aborts the communicator (ncclCommAbort) so that the other ranks fail fast instead
 * of hanging in a collective; sk_nccl_abort does it explicitly */
int sk_nccl_abort(void);
int sk_nccl_destroy(void);

/* ---- data parallel over NVLink peer memory (no collective library on the data path) ----------------
 * new ("the gradient all-reduce is the only exchange step"; the reference has no
 * multi-GPU code to replace).  Per gradient bucket: copy engines pull this rank's shard of the peers'
 * gradients, one local kernel does the sum + Adam on the shard + the GEMM weights' fp16x3 operand split,
 * copy engines push the new weights into every replica (csrc/dp_p2p.cu).
 * The arenas (one cudaMalloc block each, same layout on every rank) are exchanged as CUDA IPC handles. */
#define SK_IPC_HANDLE_BYTES 64
#define SK_P2P_MAX_WORLD 8
#define SK_P2P_MAX_BUCKETS 256
int sk_ipc_export(const void *ptr, char handle[SK_IPC_HANDLE_BYTES], int64_t *offset);
int sk_ipc_open(const char handle[SK_IPC_HANDLE_BYTES], int64_t offset, void **ptr);
int sk_ipc_close_all(void);
typedef struct {
    int world, rank, n_buckets, n_slots;
    float *grads[SK_P2P_MAX_WORLD];          /* fp32 gradient arenas, [rank] = this process's own */
    float *params[SK_P2P_MAX_WORLD];         /* fp32 parameter arenas */
    void *hi[SK_P2P_MAX_WORLD];              /* fp16 operand-split arenas (element offsets as in params), or NULL */
    void *lo[SK_P2P_MAX_WORLD];
    /* uint32 words: ready[n_buckets][8], done[n_buckets][8], parts[2][n_slots][8] (zero at start, except
     * parts[1][slot][*] = bit pattern of max |w| of every GEMM weight before the first step) */
    unsigned int *flags[SK_P2P_MAX_WORLD];
} sk_p2p_peers;
typedef struct {
    int64_t offset;          /* element offset of the tensor in the arenas */
    int64_t start, count;    /* the part of it inside this rank's piece of the bucket (count may be 0) */
    float *m, *v;            /* Adam moments of the shard (local, count elements) */
    float *scale4;           /* float[4] of the weight's operand split on THIS rank, or NULL (no split) */
    int slot;                /* row of the parts table (one per tensor) */
    int first;               /* first update of this tensor (optim.pyx:224-238) */
} sk_p2p_tensor;
typedef struct {
    double lr, beta1, beta2, eps, weight_decay, one_minus_beta1_t, one_minus_beta2_t, grad_scale, update_bound;
    int share_grads;         /* also leave the reduced gradient (sum over ranks) in every replica's arena */
    int lazy_master;         /* GEMM weights reach the replicas as hi / lo only; fp32 copies via sk_dp_p2p_gather */
} sk_p2p_adam;
/* One bucket = arena elements [bucket_start, bucket_start + bucket_len); rank r owns the piece
 * [bucket_start + r * L, bucket_start + min((r + 1) * L, bucket_len)) with L = sk_p2p_shard_len(bucket_len, world).
 * On the current launch stream (sk_launch_stream), after the bucket's last gradient kernel in stream order:
 * ready flags -> copy engines pull the peers' gradient pieces into `staging` (world * L floats, local) -> one
 * local kernel (sum in rank order, Adam, operand split) -> copy engines push the piece of params / hi / lo
 * (/ grads with share_grads) into every replica -> done flags.  `tensors` = the bucket's tensors intersected
 * with this rank's piece (count 0 where empty); scratch = SK_P2P_MAX_BUCKETS + n_slots zeroed device words. */
int64_t sk_p2p_shard_len(int64_t bucket_len, int world);
int sk_dp_p2p_update(const sk_p2p_peers *peers, int bucket, unsigned int step, int64_t bucket_start, int64_t bucket_len,
                     float *staging, int n_tensors, const sk_p2p_tensor *tensors, const sk_p2p_adam *hyper,
                     unsigned int *scratch);
/* push this rank's piece of the fp32 parameters of a bucket into every replica (current launch stream) */
int sk_dp_p2p_gather(const sk_p2p_peers *peers, int64_t bucket_start, int64_t bucket_len);
/* copy-engine probe between two device buffers (either may be a peer mapping from sk_ipc_open): reps x n_copies
 * cudaMemcpyAsync of `bytes` each, round-robin over n_streams streams; *ms = CUDA-event time of the batch */
int sk_p2p_copy_probe(void *dst, const void *src, size_t bytes, int n_copies, int n_streams, int reps, float *ms);
/* the current launch stream waits until every peer has finished `step` on the buckets in the mask */
int sk_dp_p2p_wait(const unsigned int *flags, int n_buckets, int world, unsigned int step, const unsigned int *bucket_mask);

#ifdef __cplusplus
}
#endif
#endif /* SOKET_B200_H */
