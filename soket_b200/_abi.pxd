# _abi.pxd -- Cython view of include/soket_b200.h (the C-ABI of libsoketb200.so).
from libc.stdint cimport int32_t, int64_t, uint32_t, uint64_t

cdef extern from "soket_b200.h" nogil:
    enum: SK_OK
    enum: SK_MAX_NDIM
    enum: SK_NCCL_ID_BYTES

    enum: SK_BOOL
    enum: SK_I8
    enum: SK_U8
    enum: SK_I16
    enum: SK_U16
    enum: SK_I32
    enum: SK_U32
    enum: SK_I64
    enum: SK_U64
    enum: SK_F16
    enum: SK_F32
    enum: SK_F64
    enum: SK_BF16

    enum: SK_OP_ADD
    enum: SK_OP_SUB
    enum: SK_OP_MUL
    enum: SK_OP_DIV
    enum: SK_OP_POW
    enum: SK_OP_MAXIMUM
    enum: SK_OP_MINIMUM
    enum: SK_OP_EQ
    enum: SK_OP_NE
    enum: SK_OP_GT
    enum: SK_OP_GE
    enum: SK_OP_LT
    enum: SK_OP_LE

    enum: SK_UOP_NEG
    enum: SK_UOP_EXP
    enum: SK_UOP_LOG
    enum: SK_UOP_SQRT
    enum: SK_UOP_RELU
    enum: SK_UOP_ABS

    enum: SK_RED_SUM
    enum: SK_RED_MEAN
    enum: SK_RED_MAX
    enum: SK_RED_MIN

    enum: SK_MM_AUTO
    enum: SK_MM_SIMT
    enum: SK_MM_TF32X3
    enum: SK_MM_TF32
    enum: SK_MM_BF16
    enum: SK_MM_F16X3

    enum: SK_EPI_NONE
    enum: SK_EPI_BIAS
    enum: SK_EPI_BIAS_RELU
    enum: SK_EPI_RELU

    ctypedef struct sk_array:
        void *data
        int32_t dtype
        int32_t ndim
        int64_t shape[8]
        int64_t strides[8]

    int sk_init(int device)
    int sk_device_count(int *count)
    int sk_current_device(int *device)
    int sk_sync()
    const char *sk_last_error()
    const char *sk_version()
    void *sk_stream()

    int sk_malloc(size_t nbytes, void **ptr)
    int sk_free(void *ptr)
    int sk_empty_cache()
    int sk_mem_stats(size_t *in_use, size_t *reserved, size_t *peak_in_use)
    int sk_host_alloc(size_t nbytes, void **ptr)
    int sk_host_free(void *ptr)
    int sk_h2d(void *dst, const void *src, size_t nbytes)
    int sk_d2h(void *dst, const void *src, size_t nbytes)
    int sk_h2d_async(void *dst, const void *src, size_t nbytes)
    int sk_d2h_async(void *dst, const void *src, size_t nbytes)
    int sk_d2d(void *dst, const void *src, size_t nbytes)
    int sk_memset(void *dst, int byte, size_t nbytes)

    int sk_h2d_prefetch(void *dst, const void *src, size_t nbytes)
    int sk_prefetch_wait()
    int sk_event_create(void **ev)
    int sk_event_record(void *ev)
    int sk_event_sync(void *ev)
    int sk_event_elapsed_ms(void *start, void *stop, float *ms)
    int sk_event_destroy(void *ev)
    int sk_launch_stream(int stream_id)
    int sk_event_record_on(void *ev, int stream_id)
    int sk_stream_wait_event(int stream_id, void *ev)
    uint64_t sk_launch_count()
    int sk_flush_l2()

    int sk_prof_enable(int on)
    int sk_prof_reset()
    int sk_prof_collect(int family, int64_t *launches, double *total_ms, double *total_work)

    int sk_arena_create(int *arena)
    int sk_arena_begin(int arena)
    int sk_arena_end()
    int sk_arena_destroy(int arena)
    int sk_rng_epoch_advance()
    int sk_graph_begin()
    int sk_graph_capturing()
    int sk_graph_end(void **graph_exec)
    int sk_graph_launch(void *graph_exec)
    int sk_graph_destroy(void *graph_exec)

    int sk_ewise_binary(int op, const sk_array *a, const sk_array *b, sk_array *out)
    int sk_ewise_scalar(int op, const sk_array *a, double fscalar, int64_t iscalar,
                        int scalar_is_int, int reverse, sk_array *out)
    int sk_ewise_unary(int op, const sk_array *a, sk_array *out)
    enum: SK_FUSED_MAX_OPS
    enum: SK_FUSED_MAX_INPUTS
    enum: SK_F_LOAD
    enum: SK_F_STORE
    enum: SK_F_BIN
    enum: SK_F_UN
    enum: SK_F_IN
    enum: SK_F_CONST
    enum: SK_F_TEMP
    enum: SK_F_FULL
    enum: SK_F_VECTOR
    enum: SK_F_SINGLE
    ctypedef struct sk_fused_program:
        int n_ops
        int n_in
        unsigned char code[48]
        unsigned char sub[48]
        unsigned char src[48]
        unsigned char idx[48]
        unsigned char rev[48]
        float cst[48]
        const float *inp "in" [8]
        int in_kind[8]
        int64_t n
        int64_t cols
    int sk_ewise_fused(const sk_fused_program *prog, float *out)
    int sk_copy(const sk_array *src, sk_array *dst)
    int sk_fill(sk_array *dst, double fvalue, int64_t ivalue, int value_is_int)
    int sk_relu_bwd(const sk_array *x, const sk_array *adj, sk_array *out)

    int sk_reduce(int op, const sk_array *inp, uint32_t axes_mask, sk_array *out)
    int sk_argreduce(int is_min, const sk_array *inp, int axis, sk_array *out)

    int sk_gather_rows(const sk_array *src, const sk_array *index, sk_array *out)
    int sk_eye(sk_array *out, int64_t k)
    int sk_one_hot(const sk_array *labels, sk_array *out)

    int sk_rng_seed(uint64_t seed)
    int sk_rng_uniform(sk_array *out, double low, double high)
    int sk_rng_normal(sk_array *out, double mean, double std)
    int sk_rng_bernoulli(sk_array *out, double p)

    int sk_matmul(const sk_array *a, const sk_array *b, sk_array *out, int algo)
    int sk_linear_fwd(const sk_array *x, const sk_array *w, const sk_array *bias,
                      sk_array *out, int epilogue, int algo)
    ctypedef struct sk_split_operand:
        const void *hi
        const void *lo
        int64_t ld
        int mn_major
        const float *scale
    int sk_split_f16(const float *x, int64_t rows, int64_t cols, int64_t ldx, const unsigned int *amax_bits,
                     void *hi, void *lo, int64_t ldh, float *scale4, float *colsum_out)
    int sk_gemm_f16x3_supported(int64_t M, int64_t N, int64_t K)
    int sk_gemm_f16x3(const sk_split_operand *a, const sk_split_operand *b, float *c, int64_t ldc, int64_t M,
                      int64_t N, int64_t K, const float *bias, int epilogue, int accumulate)
    ctypedef struct sk_ln_extras:
        void *split_hi
        void *split_lo
        float *split_scale
        const float *residual_scale
        unsigned int *dx_absmax
    int sk_layernorm_fwd_ex(const float *x, const float *gamma, const float *beta, const float *residual,
                            float *y, float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                            int relu, const sk_ln_extras *extras)
    int sk_layernorm_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                            const float *mean, const float *rstd, const float *y_out, int mask_mode,
                            float *dx, float *dgamma, float *dbeta, float *dresidual, int64_t rows,
                            int64_t cols, const sk_ln_extras *extras)
    int sk_layernorm_dropout_fwd_ex(const float *x, const float *gamma, const float *beta, float *y, float *mean,
                                    float *rstd, int64_t rows, int64_t cols, float eps, int relu, float keep,
                                    uint64_t *seed_out, const sk_ln_extras *extras)
    int sk_layernorm_dropout_bwd_ex(const float *adj, const float *x, const float *gamma, const float *beta,
                                    const float *mean, const float *rstd, int relu, float keep, float r_keep,
                                    uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                                    int64_t cols, const sk_ln_extras *extras)
    int sk_linear_bwd(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw)
    int sk_cast_bf16(const sk_array *src, sk_array *dst)
    int sk_linear_bwd_bias(const sk_array *adj, const sk_array *x, const sk_array *w, sk_array *dx, sk_array *dw, float *db)

    int sk_layernorm_fwd(const float *x, const float *gamma, const float *beta,
                         const float *residual, float *y, float *mean, float *rstd,
                         int64_t rows, int64_t cols, float eps, int relu)
    int sk_layernorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                         const float *mean, const float *rstd, const float *y_out, int mask_mode,
                         float *dx, float *dgamma, float *dbeta, float *dresidual,
                         int64_t rows, int64_t cols)
    int sk_batchnorm_fwd(const float *x, const float *gamma, const float *beta,
                         float *y, float *mean, float *rstd, float *running_mean,
                         float *running_var, int64_t rows, int64_t cols, float eps,
                         float momentum, int relu)
    int sk_batchnorm_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                         const float *mean, const float *rstd, const float *y_out, int mask_mode,
                         float *dx, float *dgamma, float *dbeta, int64_t rows, int64_t cols)
    int sk_softmax_ce_fwd_bwd(const float *logits, const void *labels, int label_dtype,
                              float *loss, float *dlogits, float *row_loss,
                              int64_t rows, int64_t classes)
    int sk_add_relu(const float *a, const float *b, float *out, int64_t n)
    int sk_dropout_fwd(const float *x, float *out, float *mask, int64_t n, float keep)
    int sk_dropout_fwd_seeded(const float *x, float *out, int64_t n, float keep, uint64_t *seed)
    int sk_layernorm_dropout_fwd(const float *x, const float *gamma, const float *beta, float *y,
                                 float *mean, float *rstd, int64_t rows, int64_t cols, float eps,
                                 int relu, float keep, uint64_t *seed)
    int sk_layernorm_dropout_bwd(const float *adj, const float *x, const float *gamma, const float *beta,
                                 const float *mean, const float *rstd, int relu, float keep, float r_keep,
                                 uint64_t seed, float *dx, float *dgamma, float *dbeta, int64_t rows,
                                 int64_t cols)
    int sk_dropout_bwd(const float *adj, float *out, int64_t n, float keep, float r_keep, uint64_t seed)
    int sk_colsum(const float *adj, const float *y_out, float *out, int64_t rows, int64_t cols)
    int sk_accumulate(float *acc, const float *part, int64_t n)

    int sk_sgd_step(int n_tensors, float *const *params, const float *const *grads,
                    const int64_t *sizes, double lr, double weight_decay, double grad_scale)
    int sk_adam_step(int n_tensors, float *const *params, const float *const *grads,
                     float *const *m, float *const *v, const int64_t *sizes, double lr,
                     double beta1, double beta2, double eps, double weight_decay,
                     double one_minus_beta1_t, double one_minus_beta2_t, int first_step,
                     double grad_scale)
    int sk_adam_step_dev(int n_tensors, float *const *params, const float *const *grads,
                         float *const *m, float *const *v, const int64_t *sizes, double lr,
                         double beta1, double beta2, double eps, double weight_decay,
                         const double *bias_state, int first_step, double grad_scale)
    int sk_adam_step_amax(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                          float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                          double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                          const double *bias_state, int first_step, double grad_scale, unsigned int *const *amax)
    ctypedef struct sk_adam_split:
        unsigned int *amax2
        void *hi
        void *lo
        float *scale4
    int sk_adam_step_split(int n_tensors, float *const *params, const float *const *grads, float *const *m,
                           float *const *v, const int64_t *sizes, double lr, double beta1, double beta2,
                           double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                           const double *bias_state, int first_step, double grad_scale,
                           const sk_adam_split *splits, double update_bound)
    int sk_adam_bias_advance(double *bias_state, double beta1, double beta2)

    int sk_nccl_available()
    int sk_nccl_unique_id(char *id)
    int sk_nccl_init(int rank, int world, const char *id)
    int sk_nccl_allreduce(float *buf, size_t count, int on_comm_stream)
    int sk_nccl_broadcast(float *buf, size_t count, int root)
    int sk_nccl_wait()
    int sk_nccl_destroy()
    int sk_nccl_allreduce_on(float *buf, size_t count, int stream_id)
    int sk_nccl_abort()

    int sk_ipc_export(const void *ptr, char *handle, int64_t *offset)
    int sk_ipc_open(const char *handle, int64_t offset, void **ptr)
    int sk_ipc_close_all()
    ctypedef struct sk_p2p_peers:
        int world
        int rank
        int n_buckets
        int n_slots
        float *grads[8]
        float *params[8]
        void *hi[8]
        void *lo[8]
        unsigned int *flags[8]
    ctypedef struct sk_p2p_tensor:
        int64_t offset
        int64_t start
        int64_t count
        float *m
        float *v
        float *scale4
        int slot
        int first
    ctypedef struct sk_p2p_adam:
        double lr
        double beta1
        double beta2
        double eps
        double weight_decay
        double one_minus_beta1_t
        double one_minus_beta2_t
        double grad_scale
        double update_bound
        int share_grads
        int lazy_master
    int sk_dp_p2p_gather(const sk_p2p_peers *peers, int64_t bucket_start, int64_t bucket_len)
    int64_t sk_p2p_shard_len(int64_t bucket_len, int world)
    int sk_dp_p2p_update(const sk_p2p_peers *peers, int bucket, unsigned int step, int64_t bucket_start, int64_t bucket_len,
                         float *staging, int n_tensors, const sk_p2p_tensor *tensors, const sk_p2p_adam *hyper,
                         unsigned int *scratch)
    int sk_dp_p2p_wait(const unsigned int *flags, int n_buckets, int world, unsigned int step, const unsigned int *bucket_mask)
    int sk_p2p_copy_probe(void *dst, const void *src, size_t nbytes, int n_copies, int n_streams, int reps, float *ms)
