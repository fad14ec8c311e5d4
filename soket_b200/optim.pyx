# cython: language_level=3, boundscheck=True, wraparound=True, cdivision=True
"""soket_b200.optim -- SGD and Adam (soket/optim.pyx) as single multi-tensor kernels.

Same constructor arguments and the same arithmetic, quirks included (Q2: SGD's
momentum branch only runs when momentum == 0, so every configuration is plain
SGD; Q3: Adam's `maximize` negates the gradient after its last use), but one
launch updates every parameter IN PLACE (the reference rebinds `p._data_` to a
fresh array per parameter: optim.pyx:131,254) and Adam's moments stay resident.
`grad_scale` folds the 1/W of data-parallel gradient averaging into the same
kernel.
"""
import numpy as np

from soket_b200.engine cimport Tensor
from soket_b200 import _core as B
from soket_b200 import _fused as F


cdef class Optimizer:
    """soket/optim.pyx:11-38."""
    cdef public list _params
    cdef public double grad_scale

    def __init__(self, params):
        self._params = list(params)
        self.grad_scale = 1.0

    def step(self):
        """One optimizer step over every parameter (soket/optim.pyx `step`)."""
        self.update(None)
        self.end_step()

    def update(self, subset=None):
        """Update the parameters at the given positions of the parameter list (all when None) with the
        CURRENT step's hyper-parameters.  Data-parallel training calls it once per gradient bucket,
        as each bucket's all-reduce lands, then `end_step()` once."""
        raise NotImplementedError()

    def end_step(self):
        """Advance per-step state (Adam's bias corrections)."""

    cdef tuple _live(self, subset=None):
        """Parameters that received a gradient (optim.pyx:100-102 skips the rest);
        anything the multi-tensor kernel cannot take goes to the slow list."""
        cdef list ps = [], gs = [], idx = []
        cdef Tensor p, g
        cdef int i
        for i in (range(len(self._params)) if subset is None else subset):
            p = <Tensor> self._params[i]
            if p._grad is not None:
                g = <Tensor> p._grad
                if not g._data.is_contiguous:
                    g._data = B.ascontiguousarray(g._data)
                if not p._data.is_contiguous:
                    p._data = B.ascontiguousarray(p._data)
                ps.append(p._data); gs.append(g._data); idx.append(i)
        return ps, gs, idx


cdef class SGD(Optimizer):
    """soket/optim.pyx:41-131."""
    cdef public object _lr, _momentum, _weight_decay
    cdef public bint _have_momentum, _have_weight_decay, _nesterov, _maximize

    def __init__(self, params, lr=0.01, momentum=0.0, dampening=0.0, weight_decay=0.0,
                 nesterov=False, maximize=False):
        Optimizer.__init__(self, params)
        self._lr = lr
        self._momentum = momentum
        self._have_momentum = (momentum == 0.0) is True     # quirk Q2, optim.pyx:72
        self._weight_decay = weight_decay
        self._have_weight_decay = (weight_decay != 0.0) is True
        self._nesterov = nesterov is True
        self._maximize = maximize is True
        if dampening != 0.0 and self._have_momentum:
            raise NotImplementedError('soket_b200.optim.SGD: dampening != 0 is not supported')

    def update(self, subset=None):
        ps, gs, idx = self._live(subset)
        if not ps:
            return
        lr = -self._lr if self._maximize else self._lr      # optim.pyx:127-131: p - lr * (-g)
        F.sgd_step(ps, gs, lr, self._weight_decay if self._have_weight_decay else 0.0, self.grad_scale)


_RATIO_SUP = {}
_FUSED_WEIGHT_SPLIT = True


def set_fused_weight_split(on):
    """Adam refreshes each Linear weight's fp16x3 operand split inside its own kernel (default on)."""
    global _FUSED_WEIGHT_SPLIT
    _FUSED_WEIGHT_SPLIT = bool(on)


def adam_ratio_bound(beta1, beta2, t=None):
    """An upper bound of |m_hat / sqrt(v_hat)| at step t (t=None: over every t) whatever the gradients
    were.  With m_t = (1-b1) sum_k b1^k g_{t-k} and v_t = (1-b2) sum_k b2^k g_{t-k}^2 (optim.pyx:224-247),
    Cauchy-Schwarz gives |m_t| <= (1-b1) sqrt(sum_{k<t} (b1^2/b2)^k) sqrt(v_t / (1-b2)); the bias
    corrections contribute sqrt(1-b2^t) / (1-b1^t).  Returns None when no finite bound is available."""
    b1, b2 = float(beta1), float(beta2)
    if not (0.0 <= b1 < 1.0 and 0.0 < b2 < 1.0):
        return None
    q = b1 * b1 / b2

    def at(n):
        s = float(n) if q == 1.0 else (1.0 - q ** n) / (1.0 - q)
        return (1.0 - b1) / np.sqrt(1.0 - b2) * np.sqrt(s) * np.sqrt(1.0 - b2 ** n) / (1.0 - b1 ** n)
    if t is not None:
        return float(at(int(t)))
    if q >= 1.0:
        return None                       # grows with t
    key = (b1, b2)
    if key not in _RATIO_SUP:
        n = np.arange(1, 200001, dtype=np.float64)
        sup = float(np.max((1.0 - b1) / np.sqrt(1.0 - b2) * np.sqrt((1.0 - q ** n) / (1.0 - q))
                           * np.sqrt(1.0 - b2 ** n) / (1.0 - b1 ** n)))
        _RATIO_SUP[key] = max(sup, (1.0 - b1) / np.sqrt(1.0 - b2) / np.sqrt(1.0 - q))   # and the t -> inf limit
    return _RATIO_SUP[key]


cdef class Adam(Optimizer):
    """soket/optim.pyx:134-269."""
    cdef public object _lr, _beta1, _beta2, _eps, _weight_decay
    cdef public bint _have_weight_decay
    cdef public object _t, _beta1_t, _beta2_t, _one_minus_beta1_t, _one_minus_beta2_t
    cdef public list _u, _v
    cdef public bint _capturable
    cdef public object _bias_dev

    def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, maximize=False,
                 capturable=False):
        """`capturable=True` (not in the reference): the running products beta1^t, beta2^t live in
        device memory and are advanced by a kernel, so `step()` may be captured in a CUDA graph
        (soket_b200.graph.StaticStep).  The update is bit-identical to the default form."""
        Optimizer.__init__(self, params)
        if len(betas) < 2:
            raise ValueError('Invalid betas!')
        for b in betas:
            if type(b) is not float:
                raise ValueError('Betas must be floats!')
        self._lr = lr
        self._beta1, self._beta2 = betas[0], betas[1]
        self._eps = eps
        self._weight_decay = weight_decay
        self._have_weight_decay = (weight_decay != 0.0) is True
        self._t = 1
        self._beta1_t = self._beta1
        self._beta2_t = self._beta2
        self._one_minus_beta1_t = 1.0 - self._beta1_t
        self._one_minus_beta2_t = 1.0 - self._beta2_t
        self._u = [None] * len(self._params)
        self._v = [None] * len(self._params)
        self._capturable = bool(capturable)
        self._bias_dev = None
        if self._capturable:
            self._bias_dev = B.array(np.array([self._beta1_t, self._beta2_t], dtype=np.float64))

    def update(self, subset=None):
        if B.is_capturing() and not self._capturable:
            raise RuntimeError("Adam.step inside a CUDA-graph capture: the bias corrections 1 - beta^t are host "
                               "scalars (optim.pyx:191-195) and would be frozen in the graph; construct the "
                               "optimizer with Adam(..., capturable=True) or capture SGD steps only")
        ps, gs, idx = self._live(subset)
        # first-step parameters (no state yet) and the rest go to separate launches:
        # optim.pyx:224-238 initialises m, v without the beta * 0 term
        fresh = [k for k, i in enumerate(idx) if self._u[i] is None]
        seen = [k for k, i in enumerate(idx) if self._u[i] is not None]
        if fresh and B.is_capturing():
            raise RuntimeError("Adam.step inside a CUDA-graph capture met a parameter without optimizer state: "
                               "its first-step form (optim.pyx:224-238) would be replayed forever; run one eager "
                               "step first (StaticStep's dry runs do)")
        for k in fresh:
            i = idx[k]
            self._u[i] = B.empty(ps[k].shape, 'float32')
            self._v[i] = B.empty(ps[k].shape, 'float32')
        wd = self._weight_decay if self._have_weight_decay else 0.0
        # |w_new - w_old| <= lr * |m_hat / (sqrt(v_hat) + eps)| <= lr * ratio bound: lets the kernel pick the
        # scale of the new weights' fp16x3 operand split before it has seen them (sk_adam_step_split)
        rb = adam_ratio_bound(self._beta1, self._beta2, None if self._capturable else self._t)
        ub = -1.0 if (rb is None or not _FUSED_WEIGHT_SPLIT) else abs(float(self._lr)) * rb * 1.0001
        for group, first in ((fresh, True), (seen, False)):
            if not group:
                continue
            F.adam_step([ps[k] for k in group], [gs[k] for k in group],
                        [self._u[idx[k]] for k in group], [self._v[idx[k]] for k in group],
                        self._lr, self._beta1, self._beta2, self._eps, wd,
                        self._one_minus_beta1_t, self._one_minus_beta2_t, first, self.grad_scale,
                        self._bias_dev, True, ub)

    def end_step(self):
        if self._capturable:
            # the device copy is what the kernels read; the host mirrors below only follow the
            # steps issued through this method (graph replays advance the device copy alone)
            F.adam_bias_advance(self._bias_dev, self._beta1, self._beta2)
        self._t += 1
        self._beta1_t *= self._beta1
        self._beta2_t *= self._beta2
        self._one_minus_beta1_t = 1.0 - self._beta1_t
        self._one_minus_beta2_t = 1.0 - self._beta2_t
