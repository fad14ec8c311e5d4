"""oracle/soket_np.py -- NumPy restatement of Soket's CPU path for the hot path.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs as the checker or
the timed CPU baseline -- never by ``soket_b200/`` (the product path).

Every function restates, call for call and in the same order, the sequence of
NumPy calls the reference issues (file:line under ``soket/`` cited per function;
the call log is SURVEY.md Appendix B).  Arithmetic lives in the third-party
dependency NumPy (``numpy>=2.3``, pyproject.toml:12; here 2.3.5 + OpenBLAS
0.3.30): pairwise fp32 summation, SIMD exp/log, OpenBLAS sgemm, NEP-50 weak
Python scalars.  The reference also sets MXCSR FTZ|DAZ at import
(soket/utils/ftz.pyx:18-24); that only matters below 1.2e-38.

PARITY PIN: the reference ships no tests / golden vectors (SURVEY.md section 4,
8c), so this restatement is pinned against the reference ITSELF, built from its
own sources by ``oracle/build_ref.py`` into ``oracle/_ref``:
``tests/test_oracle.py`` checks every function here against ``oracle/_ref``
bit-for-bit when it is present, and ``tests/golden/*.npz`` (generated from
``oracle/_ref`` by ``tests/golden/make_golden.py``) pins it where the built
reference cannot travel.
"""
from __future__ import annotations

import math

import numpy as np

F32 = "float32"

# Sensitivity probes (tests only).  A trajectory can only be compared as tightly as
# the CPU path agrees with ITSELF under an equally valid fp32 evaluation:
#   "hp"    every matmul evaluated in float64 and rounded to float32;
#   "splitk" the same fp32 sgemm evaluated as two half-K products added in fp32 -- a
#           different, real, summation order (no error model involved).
_MM_MODE = None


class matmul_mode:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        global _MM_MODE
        self._old = _MM_MODE
        _MM_MODE = self.mode

    def __exit__(self, *exc):
        global _MM_MODE
        _MM_MODE = self._old
        return False


def high_precision_matmul():
    return matmul_mode("hp")


def _mm(a, b, **kw):
    if _MM_MODE is None:
        return np.matmul(a, b, **kw)
    if _MM_MODE == "splitk":
        h = a.shape[-1] // 2
        if h == 0:
            return np.matmul(a, b, **kw)
        return np.add(np.matmul(a[..., :h], b[..., :h, :], dtype=F32),
                      np.matmul(a[..., h:], b[..., h:, :], dtype=F32), dtype=F32)
    return np.matmul(a.astype(np.float64), b.astype(np.float64)).astype(F32)


# ---------------------------------------------------------------- elementwise / linear
def linear_fwd(X, W, b=None):
    """Linear._fast_forward, soket/nn/prototypes.pyx:108-115 -> _matmul_fwd
    (forward.pyx:172-178) then _elemwise_add_fwd (forward.pyx:7-13)."""
    Y = _mm(X, W, dtype=F32)
    if b is not None:
        Y = np.add(Y, b, dtype=F32)
    return Y


def matmul_bwd(adj, X, W, need_dx=True):
    """_matmul_bwd, soket/tensor/ops/backward.pyx:704-742: adj @ y.T, x.T @ adj."""
    dX = _mm(adj, W.T) if need_dx else None
    dW = _mm(X.T, adj)
    return dX, dW


def broadcast_grad(adj, shape):
    """_make_gradient_compatible, soket/autodiff.pyx:43-101: a gradient that was
    broadcast (bias (H,) -> (B,H)) is summed back over the broadcast axes."""
    adj = np.asarray(adj)
    shape = tuple(shape)
    if adj.shape == shape:
        return adj
    diff = adj.ndim - len(shape)
    axes = tuple(i for i in range(adj.ndim) if i < diff or adj.shape[i] != shape[i - diff])
    data = np.sum(adj, axes, None, None, True)
    if adj.ndim != len(shape):
        data = np.reshape(data, shape)
    return data


def relu_fwd(x):
    """_relu_fwd, forward.pyx:206-209."""
    return np.maximum(x, 0)


def relu_bwd(x, adj):
    """_relu_bwd, backward.pyx:849-874: (x > 0) * adj, bool x float32."""
    return np.multiply(np.greater(x, 0), adj, dtype=F32)


def add_bwd(adj):
    """_elemwise_add_bwd, backward.pyx:60-86: each input receives a COPY of adj."""
    return np.array(adj, F32), np.array(adj, F32)


def dropout_fwd(x, mask, keep):
    """Dropout._fast_forward, prototypes.pyx:746-760: X * mask * (1/keep)."""
    r_keep = 1.0 / keep
    return np.multiply(np.multiply(x, mask, dtype=F32), r_keep, dtype=F32)


# ---------------------------------------------------------------- normalisation
def norm_fwd(Z, gamma, beta, reduce_axes, eps=1e-5, layernorm=True,
             running_mean=None, running_var=None, momentum=0.1):
    """_bnorm_fwd, forward.pyx:274-353 (training mode: quirk Q4 makes it the only
    mode).  Returns (out, xshift, rvar, norm, new_running_mean, new_running_var)."""
    mean = np.mean(Z, reduce_axes, None, None, True)
    xshift = np.subtract(Z, mean)
    var = np.mean(np.power(xshift, 2), reduce_axes, None, None, True)
    new_rm, new_rv = running_mean, running_var
    if running_mean is not None:
        sub_momentum = 1.0 - momentum
        new_rm = np.add(np.multiply(running_mean, sub_momentum), np.multiply(mean, momentum))
        new_rv = np.add(np.multiply(running_var, sub_momentum), np.multiply(var, momentum))
    rvar = np.power(np.add(var, eps), -0.5)
    norm = np.multiply(xshift, rvar)
    out = norm
    if gamma is not None:
        if layernorm:
            out = np.multiply(gamma, norm)
        else:
            out = np.multiply(np.reshape(gamma, mean.shape), norm)
        if beta is not None:
            if layernorm:
                out = np.add(beta, out)
            else:
                out = np.add(np.reshape(beta, mean.shape), out)
    return out, xshift, rvar, norm, new_rm, new_rv


def norm_bwd(adj, gamma, xshift, rvar, norm, reduce_axes, observations, layernorm=True):
    """_bnorm_bwd, backward.pyx:1025-1132.  Returns (dZ, dgamma, dbeta).
    Quirk kept: for LayerNorm dgamma/dbeta reduce over axis 0 (:1052-1055)."""
    xy_axes = (0,) if layernorm else reduce_axes
    dgamma = dbeta = None
    if gamma is not None:
        dgamma = np.sum(np.multiply(norm, adj), xy_axes, None, None, False)
        dbeta = np.sum(adj, xy_axes, None, None, False)
    robserv = 1.0 / observations
    if gamma is None:
        dxnorm = adj
    elif layernorm:
        dxnorm = np.multiply(adj, gamma)
    else:
        dxnorm = np.multiply(adj, np.reshape(gamma, rvar.shape))
    dvar = np.sum(
        np.multiply(np.multiply(dxnorm, xshift),
                    np.multiply(-0.5, np.multiply(np.multiply(rvar, rvar), rvar))),
        reduce_axes, None, None, True)
    dmean = np.add(
        np.sum(np.multiply(dxnorm, np.multiply(-1.0, rvar)), reduce_axes, None, None, True),
        np.multiply(dvar, np.multiply(robserv, np.sum(np.multiply(-2.0, xshift), reduce_axes, None, None, True))))
    dZ = np.add(np.multiply(robserv, dmean),
                np.add(np.multiply(dxnorm, rvar),
                       np.multiply(np.multiply(dvar, (2.0 * robserv)), xshift)))
    return dZ, dgamma, dbeta


# ---------------------------------------------------------------- loss
def one_hot(labels, num_classes, dtype=F32):
    """Device._one_hot, soket/backend/device.pyx:236-239: eye(C, None, 0, dtype)[labels]."""
    return np.eye(num_classes, None, 0, dtype).__getitem__(labels)


def logsumexp(x, axes, keepdims=False):
    """_logsumexp_fwd, forward.pyx:224-247."""
    m = np.max(x, axes, None, True)
    res = np.add(np.log(np.sum(np.exp(np.subtract(x, m)), axes, None, None, True)), m)
    if keepdims is not True:
        res = np.squeeze(res, axes)
    return res


def sxent_fwd(x, onehot, axes=(1,), reduction="mean"):
    """_sxentropyloss_fwd, forward.pyx:250-271."""
    batch = np.subtract(logsumexp(x, axes, False),
                        np.sum(np.multiply(x, onehot, dtype=F32), axes))
    if reduction == "sum":
        return np.sum(batch, (0,))
    if reduction == "mean":
        return np.mean(batch, (0,))
    return batch


def sxent_bwd(adj, x, onehot, axes=(1,), reduction="mean"):
    """_sxentropyloss_bwd, backward.pyx:959-1022 ('mean' / 'sum', 2-D logits)."""
    m = np.max(x, axes, None, True)
    exp_x = np.exp(np.subtract(x, m))
    sum_exp_x = np.sum(exp_x, axes, None, None, True)
    batch_grad = np.subtract(np.divide(exp_x, sum_exp_x), onehot, dtype=F32)
    if reduction == "mean" and x.ndim >= 2:
        reciprocal_batch_size = 1 / float(x.shape[0])
        batch_grad = np.multiply(batch_grad, reciprocal_batch_size)
    return np.multiply(adj, batch_grad)


def accuracy(Z, y):
    """mlp_resnet_get_accuracy, examples/mlp_resnet/model.py:61-69."""
    e = np.exp(np.subtract(Z, np.max(Z, (-1,), None, True), dtype=F32))
    softmax = np.divide(e, np.sum(e, (-1,), F32, None, True), dtype=F32)
    pred = np.array(np.argmax(softmax, -1, keepdims=False), "int32")
    return np.mean(np.equal(pred, y), None, F32, None, False).item()


# ---------------------------------------------------------------- optimisers
class SGD:
    """SGD.step, soket/optim.pyx:82-131 (quirk Q2: `_have_momentum = (momentum ==
    0.0)`, so the momentum arithmetic runs exactly when it is a no-op)."""

    def __init__(self, n_params, lr=0.01, momentum=0.0, dampening=0.0, weight_decay=0.0,
                 nesterov=False, maximize=False):
        self.lr = lr
        self.momentum = momentum
        self.have_momentum = (momentum == 0.0) is True
        self.one_minus_dampening = 1.0 - dampening
        self.weight_decay = weight_decay
        self.have_weight_decay = (weight_decay != 0.0) is True
        self.nesterov = nesterov is True
        self.maximize = maximize is True
        self.u = [None] * n_params

    def step(self, params, grads):
        out = []
        for i, (p, grad) in enumerate(zip(params, grads)):
            if grad is None:
                out.append(p)
                continue
            u = self.u[i]
            if self.have_weight_decay:
                grad = np.add(grad, np.multiply(p, self.weight_decay))
            if self.have_momentum:
                if u is None:
                    u = grad
                else:
                    u = np.add(np.multiply(u, self.momentum), np.multiply(self.one_minus_dampening, grad))
                self.u[i] = u
                if self.nesterov:
                    grad = np.add(grad, np.multiply(self.momentum, u))
                else:
                    grad = u
            if self.maximize:
                grad = np.negative(grad)
            out.append(np.subtract(p, np.multiply(self.lr, grad)))
        return out


class Adam:
    """Adam.step, soket/optim.pyx:201-269 (quirk Q3: `maximize` negates grad after
    its last use; bias corrections are Python floats)."""

    def __init__(self, n_params, lr=0.001, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, maximize=False):
        self.lr = lr
        self.beta1, self.beta2 = betas
        self.omb1 = 1.0 - self.beta1
        self.omb2 = 1.0 - self.beta2
        self.eps = eps
        self.weight_decay = weight_decay
        self.have_weight_decay = (weight_decay != 0.0) is True
        self.t = 1
        self.beta1_t = self.beta1
        self.beta2_t = self.beta2
        self.omb1_t = 1.0 - self.beta1_t
        self.omb2_t = 1.0 - self.beta2_t
        self.u = [None] * n_params
        self.v = [None] * n_params

    def step(self, params, grads):
        out = []
        for i, (p, grad) in enumerate(zip(params, grads)):
            if grad is None:
                out.append(p)
                continue
            u, v = self.u[i], self.v[i]
            if self.have_weight_decay:
                grad = np.add(grad, np.multiply(p, self.weight_decay))
            if u is None:
                u = np.multiply(grad, self.omb1)
            else:
                u = np.add(np.multiply(u, self.beta1), np.multiply(grad, self.omb1))
            self.u[i] = u
            if v is None:
                v = np.multiply(grad, np.multiply(grad, self.omb2))
            else:
                v = np.add(np.multiply(v, self.beta2), np.multiply(grad, np.multiply(grad, self.omb2)))
            self.v[i] = v
            u = np.divide(u, self.omb1_t)
            v = np.divide(v, self.omb2_t)
            out.append(np.subtract(p, np.multiply(self.lr, np.divide(u, np.add(np.power(v, 0.5), self.eps)))))
        self.t += 1
        self.beta1_t *= self.beta1
        self.beta2_t *= self.beta2
        self.omb1_t = 1.0 - self.beta1_t
        self.omb2_t = 1.0 - self.beta2_t
        return out


# ---------------------------------------------------------------- init
def kaiming_normal_std(shape, nonlinearity="relu"):
    """soket/nn/init.py:33-69 (quirk Q9): the value handed to randn as `std` is
    gain**2 / fan_in, i.e. the VARIANCE."""
    fan_in = shape[-2]
    gain = 1.4142  # soket/nn/init.py:47
    return gain * gain / fan_in


# ---------------------------------------------------------------- MLPResNet (examples/mlp_resnet/model.py)
class MLPResNet:
    """MLPResNet(dim, hidden, num_blocks, num_classes, norm) of
    examples/mlp_resnet/model.py:17-58 with explicit forward / backward in the
    order Soket's autodiff visits the graph (soket/autodiff.pyx:106-164).

    Parameter layout: ``params`` is a flat dict name -> array:
      lin0.W (dim,h) lin0.b (h,)
      blk{i}.lin1.W/b, blk{i}.n1.g/b, blk{i}.lin2.W/b, blk{i}.n2.g/b
      out.W (h,C) out.b (C,)
    ``norm`` is 'layer' or 'batch'.  Dropout is the identity (p = 0 parity runs:
    binomial(1, 1.0) is all ones, prototypes.pyx:751-758).
    """

    def __init__(self, dim, hidden, num_blocks, num_classes, norm="layer", eps=1e-5, momentum=0.1):
        self.dim, self.hidden, self.num_blocks, self.num_classes = dim, hidden, num_blocks, num_classes
        self.norm, self.eps, self.momentum = norm, eps, momentum
        self.params = {}
        z = lambda *s: np.zeros(s, F32)
        o = lambda *s: np.ones(s, F32)
        self.params["lin0.W"], self.params["lin0.b"] = z(dim, hidden), z(hidden)
        for i in range(num_blocks):
            for j in (1, 2):
                self.params[f"blk{i}.lin{j}.W"], self.params[f"blk{i}.lin{j}.b"] = z(hidden, hidden), z(hidden)
                self.params[f"blk{i}.n{j}.g"], self.params[f"blk{i}.n{j}.b"] = o(hidden), z(hidden)
        self.params["out.W"], self.params["out.b"] = z(hidden, num_classes), z(num_classes)
        # BatchNorm running stats start as 0-d tensors (prototypes.pyx:550-551)
        self.running = {}
        if norm == "batch":
            for i in range(num_blocks):
                for j in (1, 2):
                    self.running[f"blk{i}.n{j}"] = (np.array(0.0, F32), np.array(1.0, F32))

    def names(self):
        return list(self.params.keys())

    def linear_weight_names(self):
        return [k for k in self.params if k.endswith(".W")]

    def init_kaiming(self, seed=0):
        """kaiming_normal on every Linear weight under np.random.seed(seed), in
        parameter order (the `self.fn`-retaining model variant of SURVEY.md Q1)."""
        np.random.seed(seed)
        for k in self.linear_weight_names():
            shp = self.params[k].shape
            self.params[k] = np.random.normal(0.0, kaiming_normal_std(shp), shp).astype(F32)

    def _norm_axes(self):
        return ((1,), self.hidden) if self.norm == "layer" else ((0,), None)

    def forward(self, X):
        P = self.params
        ln = self.norm == "layer"
        axes = (1,) if ln else (0,)
        tape = {"X": X}
        h = linear_fwd(X, P["lin0.W"], P["lin0.b"])
        tape["lin0.pre"] = h
        h = relu_fwd(h)
        for i in range(self.num_blocks):
            blk = {"in": h}
            a = linear_fwd(h, P[f"blk{i}.lin1.W"], P[f"blk{i}.lin1.b"])
            blk["n1.in"] = a
            rm, rv = self.running.get(f"blk{i}.n1", (None, None))
            a, blk["n1.xs"], blk["n1.r"], blk["n1.norm"], rm, rv = norm_fwd(
                a, P[f"blk{i}.n1.g"], P[f"blk{i}.n1.b"], axes, self.eps, ln, rm, rv, self.momentum)
            if not ln:
                self.running[f"blk{i}.n1"] = (rm, rv)
            blk["relu1.in"] = a
            a = relu_fwd(a)
            blk["lin2.in"] = a
            a = linear_fwd(a, P[f"blk{i}.lin2.W"], P[f"blk{i}.lin2.b"])
            rm, rv = self.running.get(f"blk{i}.n2", (None, None))
            a, blk["n2.xs"], blk["n2.r"], blk["n2.norm"], rm, rv = norm_fwd(
                a, P[f"blk{i}.n2.g"], P[f"blk{i}.n2.b"], axes, self.eps, ln, rm, rv, self.momentum)
            if not ln:
                self.running[f"blk{i}.n2"] = (rm, rv)
            s = np.add(h, a, dtype=F32)  # Residual: X + fn(X), prototypes.pyx:272-273
            blk["relu2.in"] = s
            h = relu_fwd(s)
            tape[f"blk{i}"] = blk
        tape["out.in"] = h
        logits = linear_fwd(h, P["out.W"], P["out.b"])
        tape["logits"] = logits
        self.tape = tape
        return logits

    def loss(self, logits, y):
        self.onehot = one_hot(y, self.num_classes)
        return sxent_fwd(logits, self.onehot)

    def backward(self, relu_masks=None):
        """Gradients of mean softmax-CE w.r.t. every parameter.

        relu_masks (tests only, default None = the reference's own `x > 0`): the 0/1 pattern to use
        for each ReLU's derivative, as [lin0, blk0.relu1, blk0.relu2, blk1.relu1, ...].  ReLU's
        derivative is discontinuous at 0, so a pre-activation within rounding distance of 0 can get
        a different mask on another backend and move whole gradient columns by O(1/batch); the
        wide-model parity test hands the DEVICE's pattern in here -- after checking that it differs
        from this one only at such rounding-level pre-activations -- to compare gradients at 1e-5
        under one and the same mask."""
        P, T = self.params, self.tape
        if relu_masks is not None:
            relu_masks = list(relu_masks)
            assert len(relu_masks) == 1 + 2 * self.num_blocks

        def relu_back(key_x, adj, slot):
            if relu_masks is None:
                return relu_bwd(key_x, adj)
            return np.multiply(relu_masks[slot], adj, dtype=F32)
        ln = self.norm == "layer"
        axes = (1,) if ln else (0,)
        G = {}
        # tests only: the adjoint each Linear's matmul received, by layer name (with the layer inputs on
        # the tape this gives the fp32 GEMM error bound 1e-5 (|in|.T @ |adj|) of every weight gradient)
        A = self.adj_mm = {}
        adj = sxent_bwd(np.ones((), F32), T["logits"], self.onehot)
        # logits = h @ out.W + out.b
        g_mm, g_b = add_bwd(adj)
        A["out"] = g_mm
        G["out.b"] = broadcast_grad(g_b, P["out.b"].shape)
        dh, G["out.W"] = matmul_bwd(g_mm, T["out.in"], P["out.W"])
        B = T["X"].shape[0]
        for i in reversed(range(self.num_blocks)):
            blk = T[f"blk{i}"]
            obs = self.hidden if ln else B
            d = relu_back(blk["relu2.in"], dh, 2 + 2 * i)
            d_res, d_fn = add_bwd(d)
            dz, G[f"blk{i}.n2.g"], G[f"blk{i}.n2.b"] = norm_bwd(
                d_fn, P[f"blk{i}.n2.g"], blk["n2.xs"], blk["n2.r"], blk["n2.norm"], axes, obs, ln)
            g_mm, g_b = add_bwd(dz)
            G[f"blk{i}.lin2.b"] = broadcast_grad(g_b, P[f"blk{i}.lin2.b"].shape)
            A[f"blk{i}.lin2"] = g_mm
            da, G[f"blk{i}.lin2.W"] = matmul_bwd(g_mm, blk["lin2.in"], P[f"blk{i}.lin2.W"])
            da = relu_back(blk["relu1.in"], da, 1 + 2 * i)
            dz, G[f"blk{i}.n1.g"], G[f"blk{i}.n1.b"] = norm_bwd(
                da, P[f"blk{i}.n1.g"], blk["n1.xs"], blk["n1.r"], blk["n1.norm"], axes, obs, ln)
            g_mm, g_b = add_bwd(dz)
            G[f"blk{i}.lin1.b"] = broadcast_grad(g_b, P[f"blk{i}.lin1.b"].shape)
            A[f"blk{i}.lin1"] = g_mm
            dx1, G[f"blk{i}.lin1.W"] = matmul_bwd(g_mm, blk["in"], P[f"blk{i}.lin1.W"])
            # block input has two partial adjoints, summed in list order
            # (autodiff.pyx:30-41): the Residual add's copy first, then Linear1's dX
            dh = np.add(d_res, dx1, dtype=F32)
        d = relu_back(T["lin0.pre"], dh, 0)
        g_mm, g_b = add_bwd(d)
        G["lin0.b"] = broadcast_grad(g_b, P["lin0.b"].shape)
        A["lin0"] = g_mm
        _, G["lin0.W"] = matmul_bwd(g_mm, T["X"], P["lin0.W"], need_dx=False)
        self.grads = G
        return G

    def train_step(self, X, y, optim, trainable=None, relu_masks=None):
        """One step: forward, loss, backward, optimiser.  Returns the loss (float).  `relu_masks` is
        either None, a list (see backward) or a callable(self) -> list evaluated after the forward."""
        logits = self.forward(X)
        loss = self.loss(logits, y)
        if callable(relu_masks):
            relu_masks = relu_masks(self)
        G = self.backward(relu_masks)
        names = self.names() if trainable is None else list(trainable)
        new = optim.step([self.params[k] for k in names], [G[k] for k in names])
        for k, v in zip(names, new):
            self.params[k] = v
        return float(loss), logits
