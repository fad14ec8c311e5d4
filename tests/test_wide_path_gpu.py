"""End-to-end parity of the BENCHMARKED kernels against the oracle (BASELINE.md section 4 row 4,
SURVEY.md section 8d-4: "loss-parity check at a reduced size (h=512, B=1024, 50 steps)").

The engine tests in test_engine_gpu.py run the example's size (hidden 100, batch 100), where every
GEMM takes the CUDA-core kernel.  Here the model is wide enough that the step is made of the same
kernels bench.py times -- the CTA-pair tcgen05 GEMM on fp16 hi/lo operand splits, the staged
LayerNorm kernels (rows >> CTAs, so rows are claimed dynamically), the shared adj split of Linear
backward with the bias gradient as its by-product, multi-tensor SGD / Adam -- composed by the
engine exactly as in the benchmark, fusion on.  `sk.profile_collect()` proves which GEMM family
ran.

Reference semantics: examples/mlp_resnet/model.py:40-58,72-95, soket/tensor/ops/forward.pyx:172-178,
backward.pyx:704-742.  Bars (north_star): 1e-5 relative per fp32 op result -- here per training
step with both sides started from identical state (teacher forcing): the loss at 1e-5, every gradient
tensor at 1e-5 in the rms sense (||err|| <= 1e-5 ||want||), and element-wise
|err| <= 2e-5 |want| + 4e-5 rms(want) for what the adjoint reaching a layer has accumulated on its way
down (16 GEMMs and 16 LayerNorm backwards for the first block) plus, for weight gradients
dW = in.T @ adj, the fp32 GEMM bound of the kernel tests, 1e-5 (|in|.T @ |adj|) (an element is a sum of
`batch` products of either sign: its rounding is set by the size of the terms; the classical bound for
a K = 1024 fp32 dot product is 6e-5 of that sum) -- and 1e-4 on the loss at the end of a
free-running run on top of what the CPU path does to itself under a 1e-7 perturbation.
"""
import numpy as np
import pytest

from oracle import ref_model, soket_np as O
from oracle.ref_model import device_relu_signs, oracle_relu_signs, reconcile_relu_masks

pytestmark = pytest.mark.gpu

DIM, HIDDEN, BLOCKS, CLASSES, BATCH, STEPS = 784, 512, 8, 10, 1024, 50


def elementwise_excess(got, want, rtol=1e-5, floor_frac=1e-5, abs_floor=0.0):
    """max over elements of |err| / (rtol |want| + floor), floor = floor_frac * rms(want): <= 1 passes.
    The floor covers elements that are small through cancellation (a sum of 1024 per-sample terms of
    either sign): their error is set by the size of the terms, not of the result."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    rms = float(np.sqrt(np.mean(want * want))) if want.size else 0.0
    bound = rtol * np.abs(want) + floor_frac * max(rms, 1e-30) + abs_floor
    return float((np.abs(got - want) / bound).max()) if want.size else 0.0


def rms_rel_err(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return float(np.sqrt(np.mean((got - want) ** 2)) / max(np.sqrt(np.mean(want * want)), 1e-30))


def make_pair(norm="layer", seed=0, hidden=HIDDEN, blocks=BLOCKS):
    import soket_b200.api as soket
    from soket_b200 import nn
    rng = np.random.default_rng(seed)
    om = O.MLPResNet(DIM, hidden, blocks, CLASSES, norm=norm)
    for k in om.params:
        shp = om.params[k].shape
        if k.endswith(".W"):
            om.params[k] = (rng.standard_normal(shp) * np.sqrt(2.0 / shp[0])).astype("float32")
        elif ".n" in k and k.endswith(".g"):
            om.params[k] = (1 + 0.1 * rng.standard_normal(shp)).astype("float32")
        else:
            om.params[k] = (0.1 * rng.standard_normal(shp)).astype("float32")
    model = ref_model.build_model(nn, DIM, hidden, blocks, CLASSES, norm=norm, drop_prob=0.0)
    named = ref_model.named_parameters(model, blocks)
    for k, t in named.items():
        t.data = soket.Tensor(om.params[k].copy())
    return om, model, named


def sync_device_from_oracle(soket, sk, named, om, dev_opt, ora_opt, opt):
    for k, t in named.items():
        t.data = soket.Tensor(om.params[k].copy())
    if opt == "adam":
        idx = {id(t): i for i, t in enumerate(dev_opt._params)}
        for j, k in enumerate(om.names()):
            i = idx[id(named[k])]
            shp = om.params[k].shape
            dev_opt._u[i] = None if ora_opt.u[j] is None else sk.array(np.ascontiguousarray(ora_opt.u[j], dtype="float32").reshape(shp))
            dev_opt._v[i] = None if ora_opt.v[j] is None else sk.array(np.ascontiguousarray(ora_opt.v[j], dtype="float32").reshape(shp))


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_wide_model_50_steps_teacher_forced(sk, opt):
    """50 training steps of MLPResNet(784, 512, 8 blocks) at batch 1024, each started from the
    oracle's state: loss within 1e-5 and gradients / updated parameters element-wise at 1e-5 on
    EVERY step.  With 9 M ReLU inputs per step, one or two land within rounding distance of zero and
    get a different 0/1 derivative on the two backends (which moves O(1/batch) of every upstream
    gradient); the device's pattern is checked to differ only at such entries and is then used for
    the oracle's backward too, so the comparison is at one and the same point of the derivative."""
    import soket_b200.api as soket
    from soket_b200 import nn
    from soket_b200.optim import SGD, Adam
    assert nn.fusion_enabled()
    om, model, named = make_pair()
    names = om.names()
    lr = 0.01 if opt == "sgd" else 0.001
    if opt == "sgd":
        oo, do = O.SGD(len(names), lr=lr), SGD(model.parameters(), lr=lr)
    else:
        oo, do = O.Adam(len(names), lr=lr), Adam(model.parameters(), lr=lr)
    crit = nn.SoftmaxCrossEntropyLoss()
    rng = np.random.default_rng(3)
    sk.profile_reset()
    total_flips = 0
    worst_loss, worst_grad, worst_param, worst_rms = 0.0, 0.0, 0.0, 0.0
    for s in range(STEPS):
        sync_device_from_oracle(soket, sk, named, om, do, oo, opt)
        X = rng.random((BATCH, DIM), dtype=np.float32)
        y = rng.integers(0, CLASSES, BATCH).astype(np.uint8)
        dev_signs = device_relu_signs(model, X, BLOCKS)
        sk.profile_enable(True)
        loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
        loss.backward()
        sk.profile_enable(False)
        grads = {k: named[k].grad.numpy() for k in names}
        do.step()
        flips = []

        def masks(o):
            m, n = reconcile_relu_masks(o, dev_signs, BLOCKS)
            flips.append(n)
            return m
        want, _ = om.train_step(X, y, oo, relu_masks=masks)
        total_flips += flips[0]
        assert flips[0] <= 64, flips
        got = loss.item()
        worst_loss = max(worst_loss, abs(got - want) / max(1.0, abs(want)))
        assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (s, got, want)
        G = om.grads
        # a bias in front of a LayerNorm has an exactly-zero gradient in real arithmetic; both sides
        # hold the rounding residue of a sum of `batch` terms there -> absolute floor from the terms
        gscale = max(float(np.abs(np.asarray(g)).max()) for g in G.values())
        for k in names:
            g = np.asarray(G[k]).reshape(grads[k].shape)
            # 1e-5 per tensor in the rms sense (exactly-zero true gradients -- biases in front of a
            # LayerNorm -- are rounding residue on both sides and are held to the absolute floor only)
            if float(np.abs(g).max()) > 1e-4 * gscale:
                rr = rms_rel_err(grads[k], g)
                worst_rms = max(worst_rms, rr)
                assert rr <= 1e-5, (s, k, rr)
            # element-wise: |err| <= PROPAGATED + LOCAL.  Propagated: the adjoint reaching this layer is the
            # composite of up to ~50 fp32 op results (16 GEMMs and 16 LayerNorm backwards deep), each at
            # 1e-5 of its own scale -> 2e-5 |want| + 4e-5 rms(want).  Local (weights): dW = in.T @ adj is a
            # sum of `batch` products of either sign, rounded at the size of the TERMS -> the kernel
            # tests' fp32 GEMM bound 1e-5 (|in|.T @ |adj|).
            bound = 2e-5 * np.abs(g) + 4e-5 * max(float(np.sqrt(np.mean(g.astype(np.float64) ** 2))), 1e-30) + 1e-7 * gscale
            if k.endswith(".W"):
                lay = k[:-2]
                xin = (om.tape["X"] if lay == "lin0" else om.tape["out.in"] if lay == "out"
                       else om.tape[lay.split(".")[0]]["in" if lay.endswith("lin1") else "lin2.in"])
                bound = bound + 1e-5 * (np.abs(xin).T @ np.abs(om.adj_mm[lay]))
            ex = float((np.abs(grads[k].astype(np.float64) - g) / bound).max())
            worst_grad = max(worst_grad, ex)
            assert ex <= 1.0, (s, k, ex)
            if opt == "sgd":
                # p' = p - lr g: bit-level given g; compare against the oracle's update of ITS gradient
                pe = elementwise_excess(named[k].numpy(), om.params[k], rtol=1e-6, floor_frac=lr * 1e-3)
                worst_param = max(worst_param, pe)
                assert pe <= 1.0, (s, k, pe)
            else:
                # Adam: |update| <= ~lr whatever the gradient; elements whose gradient is rounding
                # residue (exactly-zero true gradients, e.g. biases in front of a LayerNorm) move
                # by up to lr in either direction on BOTH backends -> checked where the gradient
                # stands clear of its own error floor
                got_p = named[k].numpy().astype(np.float64)
                ref_p = om.params[k].astype(np.float64)
                gg = np.abs(g.astype(np.float64))
                rms = float(np.sqrt(np.mean(gg * gg)))
                clear = gg > 1e-2 * rms
                if clear.any():
                    # d(update)/d(g) ~ lr/|g| on the first steps: 1e-5 relative gradient error -> 1e-5 lr
                    assert np.abs(got_p - ref_p)[clear].max() <= 2e-3 * lr + 1e-6 * np.abs(ref_p).max(), (s, k)
                assert np.abs(got_p - ref_p).max() <= 2.0 * lr + 1e-6 * np.abs(ref_p).max(), (s, k)
    prof = sk.profile_collect()
    print(f"[wide/{opt}] {STEPS} steps, {total_flips} ReLU inputs within rounding distance of 0 in total, "
          f"worst loss rel {worst_loss:.2e}, worst grad rms rel {worst_rms:.2e}, worst grad excess {worst_grad:.3f}, worst param excess {worst_param:.3f}, "
          "families " + ", ".join(f"{k}:{v['launches']}" for k, v in prof.items()))
    # 17 forward + 33 backward GEMMs per step; only the 10-class layer may leave the tcgen05 path
    assert prof.get("gemm_tc", {}).get("launches", 0) >= STEPS * (2 * BLOCKS * 3), prof
    assert prof.get("gemm_simt", {}).get("launches", 0) <= STEPS * 3, prof


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_wide_model_50_steps_free_running(sk, opt):
    """The same 50 steps free-running: loss at the end within 1e-4 (north_star: end-of-epoch loss),
    every step within 1e-4 plus the spread of the CPU path against itself (weights perturbed by
    1e-7 relative: see test_engine_gpu.test_free_running_trajectory_layernorm).

    Learning rates: at the teacher-forced test's 0.01 / 0.001 this model (He-initialised, 8 blocks,
    random labels) overshoots -- the CPU loss goes 5.4 -> 15.9 in five steps -- and the CPU path run
    against itself with a 1e-7 weight perturbation is 0.57 (SGD) / 2e-3 (Adam) apart after 50 steps,
    so nothing is defined to 1e-4 there.  At 1e-3 (SGD) / 1e-4 (Adam) the loss falls monotonically and
    the CPU-vs-CPU spread after 50 steps is 7e-5 / 5e-4 (measured here, scripts in the commit
    message); the device has to stay inside 1e-4 + 3x that spread on every step."""
    import soket_b200.api as soket
    from soket_b200 import nn
    from soket_b200.optim import SGD, Adam
    om, model, named = make_pair()
    names = om.names()
    lr = 1e-3 if opt == "sgd" else 1e-4
    mk = (lambda: O.SGD(len(names), lr=lr)) if opt == "sgd" else (lambda: O.Adam(len(names), lr=lr))
    ens = []
    for seed in (1, 2):
        o2, _, _ = make_pair()
        r2 = np.random.default_rng(seed)
        for k in o2.params:
            o2.params[k] = (o2.params[k] * (1 + 1e-7 * r2.standard_normal(o2.params[k].shape))).astype("float32")
        ens.append((o2, mk()))
    oo = mk()
    do = SGD(model.parameters(), lr=lr) if opt == "sgd" else Adam(model.parameters(), lr=lr)
    crit = nn.SoftmaxCrossEntropyLoss()
    rng = np.random.default_rng(3)
    got, want, pert = [], [], []
    for s in range(STEPS):
        X = rng.random((BATCH, DIM), dtype=np.float32)
        y = rng.integers(0, CLASSES, BATCH).astype(np.uint8)
        loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
        loss.backward()
        do.step()
        got.append(loss.item())
        want.append(om.train_step(X, y, oo)[0])
        pert.append([o2.train_step(X, y, op2)[0] for o2, op2 in ens])
    got, want, pert = np.array(got), np.array(want), np.array(pert)
    spread = np.maximum.accumulate(np.abs(pert - want[:, None]).max(axis=1))
    err = np.abs(got - want)
    print(f"[wide-free/{opt}] final loss {got[-1]:.6f} vs {want[-1]:.6f}; max err {err.max():.2e}; CPU spread {spread[-1]:.2e}")
    assert err[0] <= 1e-5 * max(1.0, abs(want[0])), err[0]          # identical state: the per-step bar
    bound = 1e-4 * np.maximum(1.0, np.abs(want)) + 3 * spread
    assert np.all(err <= bound), (err, spread)
    assert err[-1] <= 1e-4 * max(1.0, abs(want[-1])) + 3 * spread[-1]


# ---- the benchmark's own kernel shapes ---------------------------------------------------------------
def _ln_case(rows, cols, seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype("float32")
    res = rng.standard_normal((rows, cols)).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    b = (rng.standard_normal(cols) * 0.1).astype("float32")
    adj = rng.standard_normal((rows, cols)).astype("float32")
    return x, res, g, b, adj


@pytest.mark.parametrize("rows,cols", [(8192, 4096), (8192, 512), (1024, 4096), (4099, 4096)])
def test_layernorm_kernels_at_benchmark_shape(sk, rows, cols):
    """LayerNorm forward / backward (+ReLU, +residual+ReLU) with far more rows than resident CTAs,
    i.e. on the dynamic row-claim path under contention (nn_fused.cu), element-wise against the
    oracle (forward.pyx:274-353, backward.pyx:1025-1132)."""
    from soket_b200 import _fused as F
    x, res, g, b, adj = _ln_case(rows, cols, rows + cols)
    ln, xs, rv, norm, _, _ = O.norm_fwd(x, g, b, (1,), 1e-5, True)
    dx, dg, db = O.norm_bwd(adj, g, xs, rv, norm, (1,), cols, True)
    xd, gd, bd, ad = sk.array(x), sk.array(g), sk.array(b), sk.array(adj)
    # plain
    y, mean, rstd = F.layernorm_fwd(xd, gd, bd, None, 1e-5, False)
    assert elementwise_excess(sk.asnumpy(y), ln) <= 1.0
    assert elementwise_excess(sk.asnumpy(rstd), rv[:, 0]) <= 1.0
    gx, gg, gb, _ = F.layernorm_bwd(ad, xd, gd, bd, mean, rstd)
    assert elementwise_excess(sk.asnumpy(gx), dx) <= 1.0
    # d(gamma), d(beta) are sums over `rows` terms: NumPy adds them sequentially in fp32 (axis 0),
    # the device as a tree; both are within sqrt(rows) eps of the exact sum -> compare with float64
    dg64 = (norm.astype(np.float64) * adj).sum(0)
    db64 = adj.astype(np.float64).sum(0)
    assert elementwise_excess(sk.asnumpy(gg), dg64, floor_frac=2e-5) <= 1.0
    assert elementwise_excess(sk.asnumpy(gb), db64, floor_frac=2e-5) <= 1.0
    # LN -> ReLU (model.py:28-30)
    y1, mean1, rstd1 = F.layernorm_fwd(xd, gd, bd, None, 1e-5, True)
    assert elementwise_excess(sk.asnumpy(y1), O.relu_fwd(ln)) <= 1.0
    d1 = O.relu_bwd(ln, adj)
    dx1, _, _ = O.norm_bwd(d1, g, xs, rv, norm, (1,), cols, True)
    gx1 = F.layernorm_bwd(ad, xd, gd, bd, mean1, rstd1, None, 1)[0]
    flips = np.abs(ln) < 1e-6
    assert flips.sum() <= rows * cols * 1e-5
    ok = ~flips.any(axis=1)                                   # rows whose masks are unambiguous
    assert elementwise_excess(sk.asnumpy(gx1)[ok], dx1[ok]) <= 1.0
    # relu(res + LN(x)) (prototypes.pyx:272-273 + model.py:34-37), mask from the saved output
    y2, mean2, rstd2 = F.layernorm_fwd(xd, gd, bd, sk.array(res), 1e-5, True)
    s = np.add(res, ln, dtype=np.float32)
    assert elementwise_excess(sk.asnumpy(y2), O.relu_fwd(s)) <= 1.0
    d2 = O.relu_bwd(s, adj)
    dx2, _, _ = O.norm_bwd(d2, g, xs, rv, norm, (1,), cols, True)
    out2 = F.layernorm_bwd(ad, xd, gd, bd, mean2, rstd2, y2, 2, True)
    ok = ~(np.abs(s) < 1e-6).any(axis=1)
    assert elementwise_excess(sk.asnumpy(out2[0])[ok], dx2[ok]) <= 1.0
    assert elementwise_excess(sk.asnumpy(out2[3])[ok], d2[ok]) <= 1.0


def test_first_layer_gemm_and_backward_at_benchmark_shape(sk):
    """(8192, 784) @ (784, 4096) + bias, ReLU (the first layer: K = 784 is not a multiple of the
    64-element K tile) and its backward dW = X.T @ adj, db = adj.sum(0), element-wise against
    float64 with the fp32-matmul bound |err| <= 1e-5 (|a| @ |b|)."""
    rng = np.random.default_rng(11)
    Bn, I, Od = 8192, 784, 4096
    x = rng.random((Bn, I), dtype=np.float32)
    w = (rng.standard_normal((I, Od)) * np.sqrt(2.0 / I)).astype("float32")
    b = (0.1 * rng.standard_normal(Od)).astype("float32")
    adj = rng.standard_normal((Bn, Od)).astype("float32")
    y = sk.asnumpy(sk.linear(sk.array(x), sk.array(w), sk.array(b), True))
    want = x.astype(np.float64) @ w + b
    bound = 1e-5 * (np.abs(x).astype(np.float64) @ np.abs(w) + np.abs(b))
    assert np.all(np.abs(y - np.maximum(want, 0)) <= bound)
    dx, dw, db = sk.linear_bwd(sk.array(adj), sk.array(x), sk.array(w), True)
    dw64 = x.T.astype(np.float64) @ adj
    assert np.all(np.abs(sk.asnumpy(dw) - dw64) <= 1e-5 * (np.abs(x.T).astype(np.float64) @ np.abs(adj)))
    dx64 = adj.astype(np.float64) @ w.T
    assert np.all(np.abs(sk.asnumpy(dx) - dx64) <= 1e-5 * (np.abs(adj).astype(np.float64) @ np.abs(w.T)))
    db64 = adj.astype(np.float64).sum(0)
    assert np.all(np.abs(sk.asnumpy(db) - db64) <= 1e-5 * np.abs(adj).astype(np.float64).sum(0))
