# cython: language_level=3, boundscheck=False, wraparound=False, cdivision=True
"""soket_b200._fused -- Cython entry points of the fused nn / optimiser kernels.

Each function replaces a sequence of backend array calls of the reference by one
or two launches (include/soket_b200.h, "fused nn kernels" / "optimizers"):
LayerNorm / BatchNorm forward+backward (forward.pyx:274-353, backward.pyx:1025-1132),
softmax cross-entropy forward+backward (forward.pyx:250-271, backward.pyx:959-1022),
Residual+ReLU (prototypes.pyx:272-273, model.py:34-37), Dropout
(prototypes.pyx:746-760), bias gradient (autodiff.pyx:43-101), in-place gradient
accumulation (autodiff.pyx:30-41), multi-tensor SGD / Adam (optim.pyx:82-269).
All operands are contiguous float32 device arrays.
"""
from libc.stdint cimport int64_t, uint64_t
from libc.stdlib cimport malloc, free

from soket_b200._abi cimport *
from soket_b200._core cimport ndarray, _new_array, _check, _fptr, _as_device, SplitMat, AbsMax, _new_split, _bind_split
from soket_b200 import _core as _B


cdef inline float *_opt(object a) except? NULL:
    if a is None:
        return NULL
    return _fptr(<ndarray> a)


cdef inline int _rows_cols(ndarray x, int64_t *rows, int64_t *cols) except -1:
    if x._ndim < 2:
        raise ValueError('expected an input of at least 2 dimensions')
    cols[0] = x._shape[x._ndim - 1]
    rows[0] = x._numel() // cols[0] if cols[0] else 0
    return 0


cdef inline bint _split_worthwhile(int64_t rows, int64_t cols):
    """Would a Linear consuming a (rows, cols) activation take the pre-split tcgen05 path?"""
    return rows >= 256 and cols >= 256 and cols % 8 == 0


def layernorm_fwd(ndarray x, gamma=None, beta=None, residual=None, double eps=1e-5, bint relu=False,
                  bint emit_split=False, residual_split=None):
    """y = [relu]( [residual +] gamma * ((x - mean) * rstd) + beta ) over the last axis.
    Returns (y, mean, rstd); mean/rstd have shape (rows,).  emit_split: the kernel also writes y as
    the fp16 hi / lo pair of the fp16x3 GEMM (attached to y as its SplitMat); with a residual this
    needs the residual's own SplitMat (for the bound of |residual|)."""
    cdef int64_t rows, cols
    _rows_cols(x, &rows, &cols)
    cdef ndarray y = _new_array(x._ndim, x._shape, SK_F32)
    cdef ndarray mean = _new_array(1, &rows, SK_F32)
    cdef ndarray rstd = _new_array(1, &rows, SK_F32)
    cdef sk_ln_extras ex
    cdef SplitMat m = None
    cdef SplitMat rm
    if emit_split and x._ndim == 2 and _split_worthwhile(rows, cols) and (residual is None or residual_split is not None):
        m = _new_split(rows, cols)
        ex.split_hi = <void *> m.hi._ptr; ex.split_lo = <void *> m.lo._ptr
        ex.split_scale = <float *> m.scale._ptr
        ex.residual_scale = NULL
        ex.dx_absmax = NULL
        if residual is not None:
            rm = <SplitMat> residual_split
            ex.residual_scale = <const float *> rm.scale._ptr
        _check(sk_layernorm_fwd_ex(_fptr(x), _opt(gamma), _opt(beta), _opt(residual), _fptr(y),
                                   _fptr(mean), _fptr(rstd), rows, cols, <float> eps, relu, &ex))
        _bind_split(m, y)
        return y, mean, rstd
    _check(sk_layernorm_fwd(_fptr(x), _opt(gamma), _opt(beta), _opt(residual), _fptr(y),
                            _fptr(mean), _fptr(rstd), rows, cols, <float> eps, relu))
    return y, mean, rstd


cdef inline ndarray _vec_out(object out, int64_t cols):
    """A caller-owned (cols,) float32 result buffer (a slot of the data-parallel gradient arena) or a
    fresh vector."""
    cdef ndarray o
    if out is None:
        return _new_array(1, &cols, SK_F32)
    o = <ndarray> out
    if o._code != SK_F32 or o._numel() != cols or not o._is_contiguous():
        raise ValueError('expected a contiguous float32 output vector of the parameter\'s length')
    o._touch()
    return o


def layernorm_bwd(ndarray adj, ndarray x, gamma, beta, ndarray mean, ndarray rstd,
                  y_out=None, int mask_mode=0, bint want_dresidual=False, bint want_params=True,
                  out_dgamma=None, out_dbeta=None, bint emit_absmax=False):
    """Returns (dx, dgamma, dbeta, dresidual).  emit_absmax: the kernel also leaves max |dx| in a
    device word attached to dx (AbsMax), which saves the adjoint's split its own |max| pass."""
    cdef int64_t rows, cols
    _rows_cols(x, &rows, &cols)
    cdef ndarray dx = _new_array(x._ndim, x._shape, SK_F32)
    cdef ndarray dg = None, db = None, dres = None
    if want_params:
        dg = _vec_out(out_dgamma, cols)
        db = _vec_out(out_dbeta, cols)
    if want_dresidual:
        dres = _new_array(x._ndim, x._shape, SK_F32)
    cdef sk_ln_extras ex
    cdef AbsMax am
    if emit_absmax and x._ndim == 2 and _split_worthwhile(rows, cols):
        am = _B.new_absmax_word()
        ex.split_hi = NULL; ex.split_lo = NULL; ex.split_scale = NULL; ex.residual_scale = NULL
        ex.dx_absmax = <unsigned int *> am.word._ptr
        _check(sk_layernorm_bwd_ex(_fptr(adj), _fptr(x), _opt(gamma), _opt(beta), _fptr(mean), _fptr(rstd),
                                   _opt(y_out), mask_mode, _fptr(dx), _opt(dg), _opt(db), _opt(dres),
                                   rows, cols, &ex))
        _B.bind_absmax(am, dx)
        return dx, dg, db, dres
    _check(sk_layernorm_bwd(_fptr(adj), _fptr(x), _opt(gamma), _opt(beta), _fptr(mean), _fptr(rstd),
                            _opt(y_out), mask_mode, _fptr(dx), _opt(dg), _opt(db), _opt(dres),
                            rows, cols))
    return dx, dg, db, dres


def layernorm_dropout_fwd(ndarray x, gamma, beta, double eps, bint relu, double keep, bint emit_split=False):
    """y = dropout([relu](LN(x))) in one pass.  Returns (y, mean, rstd, seed); emit_split as in
    layernorm_fwd."""
    cdef int64_t rows, cols
    _rows_cols(x, &rows, &cols)
    cdef ndarray y = _new_array(x._ndim, x._shape, SK_F32)
    cdef ndarray mean = _new_array(1, &rows, SK_F32)
    cdef ndarray rstd = _new_array(1, &rows, SK_F32)
    cdef uint64_t seed = 0
    cdef sk_ln_extras ex
    cdef SplitMat m
    if emit_split and x._ndim == 2 and _split_worthwhile(rows, cols):
        m = _new_split(rows, cols)
        ex.split_hi = <void *> m.hi._ptr; ex.split_lo = <void *> m.lo._ptr
        ex.split_scale = <float *> m.scale._ptr
        ex.residual_scale = NULL
        ex.dx_absmax = NULL
        _check(sk_layernorm_dropout_fwd_ex(_fptr(x), _opt(gamma), _opt(beta), _fptr(y), _fptr(mean), _fptr(rstd),
                                           rows, cols, <float> eps, relu, <float> keep, &seed, &ex))
        _bind_split(m, y)
        return y, mean, rstd, seed
    _check(sk_layernorm_dropout_fwd(_fptr(x), _opt(gamma), _opt(beta), _fptr(y), _fptr(mean), _fptr(rstd),
                                    rows, cols, <float> eps, relu, <float> keep, &seed))
    return y, mean, rstd, seed


def layernorm_dropout_bwd(ndarray adj, ndarray x, gamma, beta, ndarray mean, ndarray rstd, bint relu,
                          double keep, double r_keep, seed, bint want_params=True,
                          out_dgamma=None, out_dbeta=None, bint emit_absmax=False):
    """Backward of layernorm_dropout_fwd from the adjoint of its output.  Returns (dx, dgamma, dbeta)."""
    cdef int64_t rows, cols
    _rows_cols(x, &rows, &cols)
    cdef ndarray dx = _new_array(x._ndim, x._shape, SK_F32)
    cdef ndarray dg = None, db = None
    if want_params:
        dg = _vec_out(out_dgamma, cols)
        db = _vec_out(out_dbeta, cols)
    cdef sk_ln_extras ex
    cdef AbsMax am
    if emit_absmax and x._ndim == 2 and _split_worthwhile(rows, cols):
        am = _B.new_absmax_word()
        ex.split_hi = NULL; ex.split_lo = NULL; ex.split_scale = NULL; ex.residual_scale = NULL
        ex.dx_absmax = <unsigned int *> am.word._ptr
        _check(sk_layernorm_dropout_bwd_ex(_fptr(adj), _fptr(x), _opt(gamma), _opt(beta), _fptr(mean), _fptr(rstd),
                                           relu, <float> keep, <float> r_keep, <uint64_t> seed, _fptr(dx),
                                           _opt(dg), _opt(db), rows, cols, &ex))
        _B.bind_absmax(am, dx)
        return dx, dg, db
    _check(sk_layernorm_dropout_bwd(_fptr(adj), _fptr(x), _opt(gamma), _opt(beta), _fptr(mean), _fptr(rstd),
                                    relu, <float> keep, <float> r_keep, <uint64_t> seed, _fptr(dx),
                                    _opt(dg), _opt(db), rows, cols))
    return dx, dg, db


def batchnorm_fwd(ndarray x, gamma=None, beta=None, running_mean=None, running_var=None,
                  double eps=1e-5, double momentum=0.1, bint relu=False):
    """Training-mode BatchNorm1d over axis 0 of (rows, cols); running stats updated in
    place when given.  Returns (y, mean, rstd) with mean/rstd of shape (cols,)."""
    if x._ndim != 2:
        raise ValueError('batchnorm_fwd: expected a 2-D input')
    cdef int64_t rows = x._shape[0], cols = x._shape[1]
    cdef ndarray y = _new_array(2, x._shape, SK_F32)
    cdef ndarray mean = _new_array(1, &cols, SK_F32)
    cdef ndarray rstd = _new_array(1, &cols, SK_F32)
    _check(sk_batchnorm_fwd(_fptr(x), _opt(gamma), _opt(beta), _fptr(y), _fptr(mean), _fptr(rstd),
                            _opt(running_mean), _opt(running_var), rows, cols, <float> eps,
                            <float> momentum, relu))
    return y, mean, rstd


def batchnorm_bwd(ndarray adj, ndarray x, gamma, beta, ndarray mean, ndarray rstd,
                  y_out=None, int mask_mode=0):
    """Returns (dx, dgamma, dbeta)."""
    if x._ndim != 2:
        raise ValueError('batchnorm_bwd: expected a 2-D input')
    cdef int64_t rows = x._shape[0], cols = x._shape[1]
    cdef ndarray dx = _new_array(2, x._shape, SK_F32)
    cdef ndarray dg = _new_array(1, &cols, SK_F32)
    cdef ndarray db = _new_array(1, &cols, SK_F32)
    _check(sk_batchnorm_bwd(_fptr(adj), _fptr(x), _opt(gamma), _opt(beta), _fptr(mean), _fptr(rstd),
                            _opt(y_out), mask_mode, _fptr(dx), _fptr(dg), _fptr(db), rows, cols))
    return dx, dg, db


def softmax_ce(ndarray logits, ndarray labels, bint want_grad=True):
    """Mean softmax cross-entropy of (rows, classes) logits against integer labels.
    Returns (loss 0-d array, dlogits or None) in ONE pass (+ a tiny mean reduction)."""
    if logits._ndim != 2 or labels._ndim != 1 or labels._shape[0] != logits._shape[0]:
        raise ValueError('softmax_ce: expected (rows, classes) logits and (rows,) labels')
    if not labels._is_contiguous():
        labels = labels._compact()
    cdef int64_t rows = logits._shape[0], classes = logits._shape[1]
    cdef ndarray loss = _new_array(0, NULL, SK_F32)
    cdef ndarray dl = None
    if want_grad:
        dl = _new_array(2, logits._shape, SK_F32)
    _check(sk_softmax_ce_fwd_bwd(_fptr(logits), <const void *> labels._ptr, labels._code,
                                 _fptr(loss), _opt(dl), NULL, rows, classes))
    return loss, dl


def add_relu(ndarray a, ndarray b):
    """relu(a + b)."""
    if a._numel() != b._numel():
        raise ValueError('add_relu: size mismatch')
    cdef ndarray out = _new_array(a._ndim, a._shape, SK_F32)
    _check(sk_add_relu(_fptr(a), _fptr(b), _fptr(out), a._numel()))
    return out


def dropout(ndarray x, double keep, bint want_mask=True):
    """(x * mask) * (1/keep), mask ~ Bernoulli(keep).  Returns (out, mask)."""
    cdef ndarray out = _new_array(x._ndim, x._shape, SK_F32)
    cdef ndarray mask = _new_array(x._ndim, x._shape, SK_F32) if want_mask else None
    _check(sk_dropout_fwd(_fptr(x), _fptr(out), _opt(mask), x._numel(), <float> keep))
    return out, mask


def dropout_seeded(ndarray x, double keep):
    """Dropout forward without a stored mask.  Returns (out, seed); see dropout_bwd."""
    cdef ndarray out = _new_array(x._ndim, x._shape, SK_F32)
    cdef uint64_t seed = 0
    _check(sk_dropout_fwd_seeded(_fptr(x), _fptr(out), x._numel(), <float> keep, &seed))
    return out, seed


def dropout_bwd(ndarray adj, double keep, double r_keep, seed):
    """(adj * r_keep) * mask with the mask regenerated from the forward's seed."""
    cdef ndarray out = _new_array(adj._ndim, adj._shape, SK_F32)
    _check(sk_dropout_bwd(_fptr(adj), _fptr(out), adj._numel(), <float> keep, <float> r_keep, <uint64_t> seed))
    return out


def colsum(ndarray adj, y_out=None):
    """Bias gradient: sum over axis 0 of (rows, cols); y_out applies a ReLU mask first."""
    cdef int64_t rows, cols
    _rows_cols(adj, &rows, &cols)
    cdef ndarray out = _new_array(1, &cols, SK_F32)
    _check(sk_colsum(_fptr(adj), _opt(y_out), _fptr(out), rows, cols))
    return out


def accumulate_(ndarray acc, ndarray part):
    """acc += part, in place."""
    if acc._numel() != part._numel():
        raise ValueError('accumulate_: size mismatch')
    acc._touch()
    _check(sk_accumulate(_fptr(acc), _fptr(part), acc._numel()))
    return acc


cdef inline bint _wants_amax(ndarray p):
    return p._ndim == 2 and p._shape[0] >= 256 and p._shape[1] >= 128 and p._code == SK_F32


cdef class _PtrLists:
    """Host-side pointer lists for the multi-tensor optimiser kernels."""
    cdef float **p
    cdef const float **g
    cdef float **m
    cdef float **v
    cdef int64_t *sizes
    cdef int n

    def __cinit__(self, int n):
        self.n = n
        self.p = <float **> malloc(max(n, 1) * sizeof(float *))
        self.g = <const float **> malloc(max(n, 1) * sizeof(float *))
        self.m = <float **> malloc(max(n, 1) * sizeof(float *))
        self.v = <float **> malloc(max(n, 1) * sizeof(float *))
        self.sizes = <int64_t *> malloc(max(n, 1) * sizeof(int64_t))
        if not (self.p and self.g and self.m and self.v and self.sizes):
            raise MemoryError()

    def __dealloc__(self):
        free(self.p); free(<void *> self.g); free(self.m); free(self.v); free(self.sizes)


def sgd_step(list params, list grads, double lr, double weight_decay=0.0, double grad_scale=1.0):
    """In-place SGD over all tensors in ONE launch (optim.pyx:82-131)."""
    cdef int n = len(params), i
    cdef _PtrLists L = _PtrLists(n)
    cdef ndarray p, g
    for i in range(n):
        p = <ndarray> params[i]; g = <ndarray> grads[i]
        if p._numel() != g._numel():
            raise ValueError(f'sgd_step: parameter {i} and its gradient differ in size')
        p._touch()
        L.p[i] = _fptr(p); L.g[i] = _fptr(g); L.sizes[i] = p._numel()
    _check(sk_sgd_step(n, L.p, L.g, L.sizes, lr, weight_decay, grad_scale))


cdef double *_bias_ptr(ndarray state) except NULL:
    if state._code != SK_F64 or state._numel() != 2 or not state._is_contiguous() or state._ptr == 0:
        raise ValueError('Adam bias state must be a contiguous float64 device array of 2 elements')
    return <double *> state._ptr


def adam_bias_advance(ndarray state, double beta1, double beta2):
    """state[0] *= beta1; state[1] *= beta2 on the device (optim.pyx:266-267)."""
    _check(sk_adam_bias_advance(_bias_ptr(state), beta1, beta2))


def adam_step(list params, list grads, list m, list v, double lr, double beta1, double beta2,
              double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
              bint first_step, double grad_scale=1.0, bias_state=None, bint emit_absmax=True,
              double update_bound=-1.0):
    """In-place Adam over all tensors in ONE launch (optim.pyx:201-269).  With `bias_state`
    (device float64 {beta1^t, beta2^t}) the kernel forms the bias corrections itself and the
    two host scalars are ignored (graph-capturable form).

    Weight matrices a Linear layer consumes through the fp16x3 GEMM (`_wants_amax`) carry two
    persistent device words {max |w|, accumulator} that the kernel keeps current; once such a weight
    also has a valid operand split (SplitMat) and `update_bound` >= 0 bounds |w_new - w_old|, the
    kernel rewrites that split in place (sk_adam_step_split) and the weight is never split by a
    pass of its own again."""
    cdef int n = len(params), i
    cdef _PtrLists L = _PtrLists(n)
    cdef ndarray p, g
    cdef list wanted = []
    cdef list fused = [None] * n          # SplitMat to refresh in the kernel
    cdef list words = [None] * n          # AbsMax (rotating) per tracked tensor
    cdef SplitMat sm
    cdef AbsMax am
    for i in range(n):
        p = <ndarray> params[i]; g = <ndarray> grads[i]
        if p._numel() != g._numel():
            raise ValueError(f'adam_step: parameter {i} and its gradient differ in size')
        if emit_absmax and _wants_amax(p):
            # looked at BEFORE p._touch(): does the derived data describe the weights as they are now?
            am = None
            if isinstance(p._meta, SplitMat):
                sm = <SplitMat> p._meta
                if sm.amax is not None and sm.valid_for(p):
                    am = <AbsMax> sm.amax
                    if (update_bound >= 0.0 and sm.ld == sm.cols and p._is_contiguous()
                            and p._numel() % 4 == 0 and (<size_t> p._ptr) % 16 == 0):
                        fused[i] = sm
            elif isinstance(p._meta, AbsMax) and (<AbsMax> p._meta).rotating and _B.has_absmax(p):
                am = <AbsMax> p._meta
            if am is None:
                am = _B.new_rotating_absmax()
            words[i] = am
            wanted.append(i)
        p._touch()
        L.p[i] = _fptr(p); L.g[i] = _fptr(g)
        L.m[i] = _fptr(<ndarray> m[i]); L.v[i] = _fptr(<ndarray> v[i])
        L.sizes[i] = p._numel()
    cdef sk_adam_split *sp = NULL
    if wanted:
        sp = <sk_adam_split *> malloc(n * sizeof(sk_adam_split))
        if sp == NULL:
            raise MemoryError()
        for i in range(n):
            sp[i].amax2 = NULL; sp[i].hi = NULL; sp[i].lo = NULL; sp[i].scale4 = NULL
        for i in wanted:
            am = <AbsMax> words[i]
            sp[i].amax2 = <unsigned int *> am.word._ptr
            if fused[i] is not None:
                sm = <SplitMat> fused[i]
                sp[i].hi = <void *> sm.hi._ptr
                sp[i].lo = <void *> sm.lo._ptr
                sp[i].scale4 = <float *> sm.scale._ptr
        try:
            _check(sk_adam_step_split(n, L.p, L.g, L.m, L.v, L.sizes, lr, beta1, beta2, eps, weight_decay,
                                      one_minus_beta1_t, one_minus_beta2_t,
                                      _bias_ptr(<ndarray> bias_state) if bias_state is not None else NULL,
                                      first_step, grad_scale, sp, update_bound if update_bound >= 0.0 else 0.0))
        finally:
            free(sp)
        for i in wanted:
            p = <ndarray> params[i]
            if fused[i] is not None:
                _bind_split(<SplitMat> fused[i], p)      # hi / lo now describe the NEW weights
            else:
                _B.bind_absmax(<AbsMax> words[i], p)
        return
    if bias_state is not None:
        _check(sk_adam_step_dev(n, L.p, L.g, L.m, L.v, L.sizes, lr, beta1, beta2, eps, weight_decay,
                                _bias_ptr(<ndarray> bias_state), first_step, grad_scale))
        return
    _check(sk_adam_step(n, L.p, L.g, L.m, L.v, L.sizes, lr, beta1, beta2, eps, weight_decay,
                        one_minus_beta1_t, one_minus_beta2_t, first_step, grad_scale))


# ---------------------------------------------------------------- data parallel
def nccl_available():
    return bool(sk_nccl_available())


def nccl_unique_id():
    cdef char buf[128]
    _check(sk_nccl_unique_id(buf))
    return bytes(buf[:128])


def nccl_init(int rank, int world, bytes uid):
    if len(uid) != 128:
        raise ValueError('nccl_init: the unique id must be 128 bytes')
    cdef const char *p = uid
    _check(sk_nccl_init(rank, world, p))


def nccl_allreduce(ndarray buf, bint on_comm_stream=False):
    """In-place sum all-reduce of a contiguous fp32 buffer."""
    buf._touch()
    _check(sk_nccl_allreduce(_fptr(buf), <size_t> buf._numel(), on_comm_stream))


def nccl_allreduce_on(ndarray buf, int stream):
    """In-place sum all-reduce on the given stream id; the caller orders it with events."""
    buf._touch()
    _check(sk_nccl_allreduce_on(_fptr(buf), <size_t> buf._numel(), stream))


def nccl_abort():
    _check(sk_nccl_abort())


def nccl_broadcast(ndarray buf, int root=0):
    buf._touch()
    _check(sk_nccl_broadcast(_fptr(buf), <size_t> buf._numel(), root))


def nccl_wait():
    _check(sk_nccl_wait())


def nccl_destroy():
    _check(sk_nccl_destroy())


# ---------------------------------------------------------------- data parallel over peer memory
def ipc_export(ndarray a):
    """(64-byte CUDA IPC handle of the allocation `a` lives in, byte offset of a's first element in it)."""
    cdef char h[64]
    cdef int64_t off = 0
    _check(sk_ipc_export(<const void *> a._ptr, h, &off))
    return bytes(h[:64]), int(off)


def ipc_open(bytes handle, int64_t offset):
    """Map a peer's allocation into this process; returns the device address of its exported array."""
    if len(handle) != 64:
        raise ValueError('ipc_open: the handle must be 64 bytes')
    cdef void *p = NULL
    _check(sk_ipc_open(<const char *> handle, offset, &p))
    return <size_t> p


def ipc_close_all():
    _check(sk_ipc_close_all())


def p2p_copy_probe(size_t dst, size_t src, size_t nbytes, int n_copies=1, int n_streams=1, int reps=1):
    """Milliseconds for reps x n_copies copy-engine copies of nbytes each between two device addresses."""
    cdef float ms = 0
    _check(sk_p2p_copy_probe(<void *> dst, <const void *> src, nbytes, n_copies, n_streams, reps, &ms))
    return ms


cdef class P2pPeers:
    """The arenas of every rank as this process sees them (soket_b200.dp, mode 'p2p')."""
    cdef sk_p2p_peers c
    cdef public ndarray scratch
    cdef public ndarray flags
    cdef public int n_buckets

    def __init__(self, int world, int rank, int n_buckets, int n_slots, list grads, list params, list hi, list lo,
                 list flags, ndarray own_flags):
        cdef int q
        if world < 2 or world > 8 or len(grads) != world or len(params) != world or len(flags) != world:
            raise ValueError('P2pPeers: 2..8 ranks, one address per rank and arena')
        self.c.world = world; self.c.rank = rank; self.c.n_buckets = n_buckets; self.c.n_slots = n_slots
        for q in range(8):
            self.c.grads[q] = NULL; self.c.params[q] = NULL; self.c.hi[q] = NULL; self.c.lo[q] = NULL; self.c.flags[q] = NULL
        for q in range(world):
            self.c.grads[q] = <float *> <size_t> grads[q]
            self.c.params[q] = <float *> <size_t> params[q]
            self.c.hi[q] = <void *> <size_t> hi[q]
            self.c.lo[q] = <void *> <size_t> lo[q]
            self.c.flags[q] = <unsigned int *> <size_t> flags[q]
        self.flags = own_flags
        self.n_buckets = n_buckets
        self.scratch = _B.zeros((256 + max(n_slots, 1),), 'uint32')


def p2p_shard_len(int64_t bucket_len, int world):
    return int(sk_p2p_shard_len(bucket_len, world))


def dp_p2p_update(P2pPeers peers, int bucket, unsigned int step, int64_t bucket_start, int64_t bucket_len,
                  ndarray staging, list tensors, double lr, double beta1, double beta2,
                  double eps, double weight_decay, double one_minus_beta1_t, double one_minus_beta2_t,
                  double grad_scale, double update_bound, bint share_grads, bint lazy_master=False):
    """One bucket of the peer-memory data-parallel Adam step (sk_dp_p2p_update) on the current launch stream.
    tensors: [(param ndarray, arena offset, start, count, m, v, SplitMat or None, slot, first)] -- start / count =
    the part of the tensor inside this rank's piece of the bucket; staging: world * shard_len float32 (local)."""
    cdef int n = len(tensors), i
    cdef sk_p2p_tensor *ts = <sk_p2p_tensor *> malloc(max(n, 1) * sizeof(sk_p2p_tensor))
    if ts == NULL:
        raise MemoryError()
    cdef sk_p2p_adam h
    cdef ndarray p
    cdef SplitMat sm
    h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay
    h.one_minus_beta1_t = one_minus_beta1_t; h.one_minus_beta2_t = one_minus_beta2_t
    h.grad_scale = grad_scale; h.update_bound = update_bound; h.share_grads = 1 if share_grads else 0
    h.lazy_master = 1 if lazy_master else 0
    try:
        for i in range(n):
            p, off, start, count, m, v, split, slot, first = tensors[i]
            ts[i].offset = off; ts[i].start = start; ts[i].count = count
            ts[i].m = _fptr(<ndarray> m) if count else NULL
            ts[i].v = _fptr(<ndarray> v) if count else NULL
            ts[i].scale4 = NULL
            if split is not None:
                sm = <SplitMat> split
                ts[i].scale4 = <float *> sm.scale._ptr
            ts[i].slot = slot; ts[i].first = 1 if first else 0
        if staging._code != SK_F32 or staging._numel() < peers.c.world * sk_p2p_shard_len(bucket_len, peers.c.world):
            raise ValueError('dp_p2p_update: staging must hold world * shard_len float32 elements')
        _check(sk_dp_p2p_update(&peers.c, bucket, step, bucket_start, bucket_len, _fptr(staging), n, ts, &h,
                                <unsigned int *> peers.scratch._ptr))
    finally:
        free(ts)
    for i in range(n):
        p = <ndarray> tensors[i][0]
        p._touch()                                       # every replica of p is rewritten by this launch
        if tensors[i][6] is not None:
            _bind_split(<SplitMat> tensors[i][6], p)     # ... and so is its operand split


def dp_p2p_gather(P2pPeers peers, int64_t bucket_start, int64_t bucket_len):
    """Push this rank's piece of a bucket's fp32 parameters into every replica (current launch stream)."""
    _check(sk_dp_p2p_gather(&peers.c, bucket_start, bucket_len))


def dp_p2p_wait(P2pPeers peers, unsigned int step, list bucket_ids):
    """The current launch stream waits until every rank has finished `step` on these buckets."""
    cdef unsigned int mask[8]
    cdef int i
    for i in range(8):
        mask[i] = 0
    for b in bucket_ids:
        mask[b >> 5] |= (<unsigned int> 1) << (b & 31)
    _check(sk_dp_p2p_wait(<const unsigned int *> peers.flags._ptr, peers.n_buckets, peers.c.world, step, mask))
