#!/usr/bin/env python
"""Generate tests/golden/*.npz from the BUILT REFERENCE (oracle/_ref).

The reference ships no golden vectors, so the pins are
outputs of the reference itself, run here on its CPU (NumPy) device with seeded
inputs.  Run where /root/reference exists:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Every array is produced by the reference's own Tensor / autodiff / nn / optim
code; nothing from oracle/soket_np.py or soket_b200 is involved.  The fixtures
are small (a few hundred KB in total) and committed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_model  # noqa: E402

soket = ref_model.import_reference()
if soket is None:
    raise SystemExit("oracle/_ref is not built")
import soket.nn as nn  # noqa: E402
from soket.optim import SGD, Adam  # noqa: E402


def to_numpy(t):
    """The reference has no accessor; Tensor.from_numpy wraps our buffer without
    copying (tensor.pyx:1073-1084) and __setitem__ writes into it (tensor.pyx:948)."""
    t = soket.Tensor(t, soket.cpu())
    if len(t.shape) == 0:
        return np.array(t.item(), dtype=str(t.dtype))
    buf = np.zeros(t.shape, dtype=str(t.dtype))
    view = soket.Tensor.from_numpy(buf)
    view[tuple(slice(None) for _ in t.shape)] = t
    return buf


def init_params(rng, dim, hidden, nb, C):
    P = {}
    def w(i, o): return (rng.standard_normal((i, o)) * np.sqrt(2.0 / i)).astype("float32")
    def b(n): return (0.1 * rng.standard_normal(n)).astype("float32")
    def g(n): return (1 + 0.1 * rng.standard_normal(n)).astype("float32")
    P["lin0.W"], P["lin0.b"] = w(dim, hidden), b(hidden)
    for i in range(nb):
        for j in (1, 2):
            P[f"blk{i}.lin{j}.W"], P[f"blk{i}.lin{j}.b"] = w(hidden, hidden), b(hidden)
            P[f"blk{i}.n{j}.g"], P[f"blk{i}.n{j}.b"] = g(hidden), b(hidden)
    P["out.W"], P["out.b"] = w(hidden, C), b(C)
    return P


def mlpresnet_case(norm, opt, dim=24, hidden=16, nb=2, C=10, B=12, steps=5, seed=0):
    rng = np.random.default_rng(seed)
    P = init_params(rng, dim, hidden, nb, C)
    model = ref_model.build_model(nn, dim, hidden, nb, C, norm=norm, drop_prob=0.0)
    named = ref_model.named_parameters(model, nb)
    for k, t in named.items():
        t.data = soket.Tensor(P[k].copy())
    if opt == "sgd":
        o = SGD(model.parameters(), lr=0.05, weight_decay=0.0)
        hyper = dict(lr=0.05, weight_decay=0.0)
    else:
        o = Adam(model.parameters(), lr=0.01, weight_decay=0.001)
        hyper = dict(lr=0.01, weight_decay=0.001)
    crit = nn.SoftmaxCrossEntropyLoss()
    X = rng.random((steps, B, dim), dtype=np.float32)
    y = rng.integers(0, C, (steps, B)).astype(np.uint8)
    out = {f"init/{k}": v for k, v in P.items()}
    out["X"], out["y"] = X, y
    out["hyper"] = np.array([hyper["lr"], hyper["weight_decay"]])
    losses = []
    for s in range(steps):
        logits = model(soket.Tensor(X[s]))
        loss = crit(logits, soket.Tensor(y[s]))
        loss.backward()
        if s == 0:
            out["logits0"] = to_numpy(logits)
            for k, t in named.items():
                out[f"grad0/{k}"] = to_numpy(t.grad)
        o.step()
        losses.append(loss.item())
    out["losses"] = np.array(losses, dtype=np.float64)
    for k, t in named.items():
        out[f"final/{k}"] = to_numpy(t)
    return out


def op_cases(seed=1):
    """Single-op vectors through the reference's Tensor API."""
    rng = np.random.default_rng(seed)
    out = {}
    a = rng.standard_normal((6, 8)).astype("float32")
    b = (rng.standard_normal((8,)) + 2).astype("float32")
    A = soket.Tensor(a, requires_grad=True)
    Bt = soket.Tensor(b, requires_grad=True)
    out["a"], out["b"] = a, b
    z = ((A * Bt + 1.5) / Bt - A ** 2)
    out["ewise"] = to_numpy(z)
    z.sum().backward()
    out["ewise_da"], out["ewise_db"] = to_numpy(A.grad), to_numpy(Bt.grad)
    out["sum_ax0"] = to_numpy(A.sum(0))
    out["mean_ax1_keep"] = to_numpy(A.mean(1, keepdims=True))
    out["max_ax1"] = to_numpy(A.max(1))
    out["argmax_ax1"] = to_numpy(A.argmax(1))
    out["transpose"] = to_numpy(A.T)
    out["reshape"] = to_numpy(A.reshape(4, 12))
    out["bcast"] = to_numpy(Bt.broadcast_to(3, 6, 8))
    out["getitem"] = to_numpy(A[1:5:2, ::3])
    out["le_quirk"] = to_numpy(A <= soket.Tensor(np.zeros((6, 8), "float32")))
    out["exp"] = to_numpy(soket.exp(A))
    out["log"] = to_numpy(soket.log(Bt))
    out["logsumexp"] = to_numpy(soket.logsumexp(A, 1))
    w = rng.standard_normal((8, 5)).astype("float32")
    W = soket.Tensor(w, requires_grad=True)
    A2 = soket.Tensor(a, requires_grad=True)
    mm = A2 @ W
    out["w"] = w
    out["matmul"] = to_numpy(mm)
    (mm * mm).sum().backward()
    out["matmul_da"], out["matmul_dw"] = to_numpy(A2.grad), to_numpy(W.grad)
    # layer norm / batch norm / loss modules
    x = (rng.standard_normal((10, 12)) * 2 + 1).astype("float32")
    for name, mod in (("ln", nn.LayerNorm(12)), ("bn", nn.BatchNorm1d(12))):
        X = soket.Tensor(x, requires_grad=True)
        g, bb = list(mod.parameters())
        gv = (1 + 0.2 * rng.standard_normal(12)).astype("float32")
        bv = (0.3 * rng.standard_normal(12)).astype("float32")
        g.data = soket.Tensor(gv); bb.data = soket.Tensor(bv)
        y = mod(X)
        coef = rng.standard_normal((10, 12)).astype("float32")
        (y * soket.Tensor(coef)).sum().backward()
        out[f"{name}_x"], out[f"{name}_g"], out[f"{name}_b"], out[f"{name}_coef"] = x, gv, bv, coef
        out[f"{name}_y"] = to_numpy(y)
        out[f"{name}_dx"], out[f"{name}_dg"], out[f"{name}_db"] = to_numpy(X.grad), to_numpy(g.grad), to_numpy(bb.grad)
    logits = (rng.standard_normal((9, 10)) * 3).astype("float32")
    labels = rng.integers(0, 10, 9).astype("uint8")
    L = soket.Tensor(logits, requires_grad=True)
    loss = nn.SoftmaxCrossEntropyLoss()(L, soket.Tensor(labels))
    loss.backward()
    out["ce_logits"], out["ce_labels"] = logits, labels
    out["ce_loss"] = to_numpy(loss)
    out["ce_dlogits"] = to_numpy(L.grad)
    # known answer: 3 SGD steps on w^2: [0.512, 1.024]
    wt = soket.Tensor([1.0, 2.0], requires_grad=True)
    o = SGD([wt], lr=0.1, momentum=0.9)
    for _ in range(3):
        (wt * wt).sum().backward()
        o.step()
    out["sgd_w2"] = to_numpy(wt)
    return out


if __name__ == "__main__":
    for norm in ("layer", "batch"):
        for opt in ("sgd", "adam"):
            np.savez_compressed(os.path.join(HERE, f"mlpresnet_{norm}_{opt}.npz"), **mlpresnet_case(norm, opt))
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **op_cases())
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
