"""GPU parity of the fused nn / optimiser kernels against the NumPy oracle
(oracle/soket_np.py, pinned bit-for-bit to the built reference by test_oracle.py).

Tolerance: 1e-5 relative (north_star) on forward values and gradients; the
multi-tensor optimisers are bit-exact given identical gradients.
"""
import numpy as np
import pytest

from oracle import soket_np as O

pytestmark = pytest.mark.gpu


def rel(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


@pytest.fixture(scope="module")
def F(sk):
    from soket_b200 import _fused
    return _fused


@pytest.mark.parametrize("rows,cols", [(100, 100), (7, 4), (33, 128), (64, 512), (200, 1024), (50, 2048),
                                       (16, 4096), (9, 8192), (1, 100), (300, 36)])
def test_layernorm_fwd_bwd(sk, F, rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    b = rng.standard_normal(cols).astype("float32")
    adj = rng.standard_normal((rows, cols)).astype("float32")
    want, xs, rv, norm, _, _ = O.norm_fwd(x, g, b, (1,), 1e-5, True)
    y, mean, rstd = F.layernorm_fwd(sk.array(x), sk.array(g), sk.array(b), None, 1e-5, False)
    assert rel(sk.asnumpy(y), want) <= 1e-5
    assert rel(sk.asnumpy(rstd), rv[:, 0]) <= 1e-5
    dz, dg, db = O.norm_bwd(adj, g, xs, rv, norm, (1,), cols, True)
    dx, dgam, dbet, _ = F.layernorm_bwd(sk.array(adj), sk.array(x), sk.array(g), sk.array(b), mean, rstd)
    assert rel(sk.asnumpy(dx), dz) <= 1e-5
    assert rel(sk.asnumpy(dgam), dg) <= 1e-5
    assert rel(sk.asnumpy(dbet), db) <= 1e-5


@pytest.mark.parametrize("rows,cols", [(100, 100), (64, 1024), (31, 4096)])
def test_layernorm_fused_relu_and_residual(sk, F, rows, cols):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((rows, cols)).astype("float32")
    res = rng.standard_normal((rows, cols)).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    b = (rng.standard_normal(cols) * 0.1).astype("float32")
    adj = rng.standard_normal((rows, cols)).astype("float32")
    ln, xs, rv, norm, _, _ = O.norm_fwd(x, g, b, (1,), 1e-5, True)
    # LN -> ReLU (model.py:28-30), mask recomputed in backward (mask_mode 1)
    y, mean, rstd = F.layernorm_fwd(sk.array(x), sk.array(g), sk.array(b), None, 1e-5, True)
    want = O.relu_fwd(ln)
    got = sk.asnumpy(y)
    assert rel(got, want) <= 1e-5
    d_ln = O.relu_bwd(ln, adj)
    # compare only where the pre-activation is not within rounding of zero
    safe = np.abs(ln) > 1e-4
    dz, dg, db = O.norm_bwd(np.where(safe, d_ln, 0).astype("float32"), g, xs, rv, norm, (1,), cols, True)
    adj_safe = np.where(safe, adj, 0).astype("float32")
    dx, dgam, dbet, _ = F.layernorm_bwd(sk.array(adj_safe), sk.array(x), sk.array(g), sk.array(b), mean, rstd, None, 1)
    assert rel(sk.asnumpy(dx), dz) <= 1e-5
    assert rel(sk.asnumpy(dgam), dg) <= 1e-5 and rel(sk.asnumpy(dbet), db) <= 1e-5
    # relu(residual + LN(x)) (prototypes.pyx:272-273 + model.py:36), mask from the output (mask_mode 2)
    y2, mean, rstd = F.layernorm_fwd(sk.array(x), sk.array(g), sk.array(b), sk.array(res), 1e-5, True)
    s = np.add(res, ln, dtype="float32")
    assert rel(sk.asnumpy(y2), O.relu_fwd(s)) <= 1e-5
    safe = np.abs(s) > 1e-4
    adj_safe = np.where(safe, adj, 0).astype("float32")
    d = O.relu_bwd(s, adj_safe)
    dz, dg, db = O.norm_bwd(d, g, xs, rv, norm, (1,), cols, True)
    dx, dgam, dbet, dres = F.layernorm_bwd(sk.array(adj_safe), sk.array(x), sk.array(g), sk.array(b), mean, rstd,
                                           y2, 2, True)
    assert rel(sk.asnumpy(dx), dz) <= 1e-5
    assert rel(sk.asnumpy(dres), d) <= 1e-6
    assert rel(sk.asnumpy(dgam), dg) <= 1e-5 and rel(sk.asnumpy(dbet), db) <= 1e-5


@pytest.mark.parametrize("rows,cols", [(100, 100), (12, 16), (2048, 512), (999, 260), (64, 4096)])
def test_batchnorm_fwd_bwd(sk, F, rows, cols):
    rng = np.random.default_rng(rows + cols)
    x = (rng.standard_normal((rows, cols)) * 1.5 + 3.0).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    b = rng.standard_normal(cols).astype("float32")
    adj = rng.standard_normal((rows, cols)).astype("float32")
    rm0 = rng.standard_normal(cols).astype("float32")
    rv0 = (rng.random(cols) + 0.5).astype("float32")
    want, xs, rv, norm, rm1, rv1 = O.norm_fwd(x, g, b, (0,), 1e-5, False, rm0.reshape(1, -1), rv0.reshape(1, -1), 0.1)
    drm, drv = sk.array(rm0), sk.array(rv0)
    y, mean, rstd = F.batchnorm_fwd(sk.array(x), sk.array(g), sk.array(b), drm, drv, 1e-5, 0.1, False)
    assert rel(sk.asnumpy(y), want) <= 1e-5
    assert rel(sk.asnumpy(drm), rm1[0]) <= 1e-5 and rel(sk.asnumpy(drv), rv1[0]) <= 1e-5
    dz, dg, db = O.norm_bwd(adj, g, xs, rv, norm, (0,), rows, False)
    dx, dgam, dbet = F.batchnorm_bwd(sk.array(adj), sk.array(x), sk.array(g), sk.array(b), mean, rstd)
    assert rel(sk.asnumpy(dx), dz) <= 1e-5
    assert rel(sk.asnumpy(dgam), dg) <= 1e-5
    assert rel(sk.asnumpy(dbet), db) <= 1e-5


@pytest.mark.parametrize("rows,classes,ldt", [(100, 10, "uint8"), (8192, 10, "uint8"), (7, 3, "int32"),
                                              (64, 1000, "int64"), (1, 10, "uint8")])
def test_softmax_ce(sk, F, rows, classes, ldt):
    rng = np.random.default_rng(rows)
    x = (rng.standard_normal((rows, classes)) * 3).astype("float32")
    y = rng.integers(0, classes, rows).astype(ldt)
    oh = O.one_hot(y, classes)
    want = O.sxent_fwd(x, oh)
    wgrad = O.sxent_bwd(np.ones((), "float32"), x, oh)
    loss, dl = F.softmax_ce(sk.array(x), sk.array(y))
    assert abs(loss.item() - float(want)) <= 1e-5 * max(abs(float(want)), 1.0)
    assert rel(sk.asnumpy(dl), wgrad) <= 1e-5


def test_add_relu_colsum_accumulate_dropout(sk, F):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((257, 100)).astype("float32")
    b = rng.standard_normal((257, 100)).astype("float32")
    assert np.array_equal(sk.asnumpy(F.add_relu(sk.array(a), sk.array(b))), O.relu_fwd(np.add(a, b, dtype="float32")))
    assert rel(sk.asnumpy(F.colsum(sk.array(a))), O.broadcast_grad(a, (100,))) <= 2e-6
    y = O.relu_fwd(b)
    assert rel(sk.asnumpy(F.colsum(sk.array(a), sk.array(y))), O.relu_bwd(b, a).sum(0)) <= 2e-6
    acc = sk.array(a)
    F.accumulate_(acc, sk.array(b))
    assert np.array_equal(sk.asnumpy(acc), np.add(a, b, dtype="float32"))
    out, mask = F.dropout(sk.array(a), 0.75)
    m = sk.asnumpy(mask)
    assert set(np.unique(m)) <= {0.0, 1.0} and abs(m.mean() - 0.75) < 0.02
    assert np.array_equal(sk.asnumpy(out), O.dropout_fwd(a, m, 0.75))
    out, mask = F.dropout(sk.array(a), 1.0)
    assert np.array_equal(sk.asnumpy(out), a)
    # mask-free variant: the backward regenerates the Bernoulli draw from the forward's seed
    keep = 0.75
    r_keep = 1.0 / keep
    out, seed = F.dropout_seeded(sk.array(a), keep)
    m = (sk.asnumpy(F.dropout_bwd(sk.ones(a.shape, "float32"), keep, 1.0, seed)))      # (1 * 1) * mask
    assert set(np.unique(m)) <= {0.0, 1.0} and abs(m.mean() - keep) < 0.02
    assert np.array_equal(sk.asnumpy(out), O.dropout_fwd(a, m, keep))
    want = np.multiply(np.multiply(b, np.float32(r_keep), dtype="float32"), m, dtype="float32")
    assert np.array_equal(sk.asnumpy(F.dropout_bwd(sk.array(b), keep, r_keep, seed)), want)
    out2, seed2 = F.dropout_seeded(sk.array(a), keep)
    assert seed2 != seed and not np.array_equal(sk.asnumpy(out2), sk.asnumpy(out))        # a fresh draw per call


@pytest.mark.parametrize("wd", [0.0, 0.001])
def test_sgd_bit_exact(sk, F, wd):
    rng = np.random.default_rng(4)
    shapes = [(784, 100), (100,), (100, 10), (10,), (3,), (4099,)]
    P = [rng.standard_normal(s).astype("float32") for s in shapes]
    dP = [sk.array(p) for p in P]
    opt = O.SGD(len(P), lr=0.01, weight_decay=wd)
    for step in range(3):
        G = [rng.standard_normal(s).astype("float32") for s in shapes]
        P = opt.step(P, G)
        F.sgd_step(dP, [sk.array(g) for g in G], 0.01, wd, 1.0)
        for p, d in zip(P, dP):
            assert np.array_equal(sk.asnumpy(d), p)


@pytest.mark.parametrize("wd", [0, 0.001])
def test_adam_bit_exact(sk, F, wd):
    rng = np.random.default_rng(5)
    shapes = [(784, 100), (100,), (100, 10), (10,), (1,), (70001,)] + [(5, 5)] * 60  # > 48 tensors: two launches
    P = [rng.standard_normal(s).astype("float32") for s in shapes]
    dP = [sk.array(p) for p in P]
    dM = [sk.zeros(s, "float32") for s in shapes]
    dV = [sk.zeros(s, "float32") for s in shapes]
    opt = O.Adam(len(P), lr=0.001, weight_decay=wd)
    for step in range(4):
        G = [(rng.standard_normal(s) * 0.1).astype("float32") for s in shapes]
        omb1_t, omb2_t = opt.omb1_t, opt.omb2_t
        P = opt.step(P, G)
        F.adam_step(dP, [sk.array(g) for g in G], dM, dV, 0.001, 0.9, 0.999, 1e-8, wd, omb1_t, omb2_t, step == 0, 1.0)
        for i, (p, d) in enumerate(zip(P, dP)):
            assert np.array_equal(sk.asnumpy(d), p), (step, i)


@pytest.mark.parametrize("rows,cols", [(64, 4096), (37, 256), (5, 1024), (300, 100)])
@pytest.mark.parametrize("relu", [True, False])
def test_layernorm_dropout_fused_equals_the_two_kernels(sk, F, rows, cols, relu):
    """sk_layernorm_dropout_fwd / _bwd (LayerNorm - ReLU - Dropout in one pass each way) against
    sk_layernorm_fwd + the seeded dropout kernels + sk_layernorm_bwd on the same draw: identical
    values (staged and register LayerNorm kernels; -0.0 == 0.0)."""
    rng = np.random.default_rng(rows * cols + relu)
    x = sk.array(rng.standard_normal((rows, cols)).astype("float32"))
    g = sk.array((1 + 0.1 * rng.standard_normal(cols)).astype("float32"))
    b = sk.array((0.1 * rng.standard_normal(cols)).astype("float32"))
    adj = sk.array(rng.standard_normal((rows, cols)).astype("float32"))
    keep = 0.7
    y, mean, rstd, seed = F.layernorm_dropout_fwd(x, g, b, 1e-5, relu, keep)
    y0, mean0, rstd0 = F.layernorm_fwd(x, g, b, None, 1e-5, relu)
    r_fwd = float(np.float32(1.0 / np.float64(np.float32(keep))))       # what the forward kernels use
    want_y = sk.asnumpy(F.dropout_bwd(y0, keep, r_fwd, seed))           # (y * r) * m == (y * m) * r
    got_y = sk.asnumpy(y)
    assert np.array_equal(got_y, want_y)
    frac = float((got_y == 0).mean())
    assert (0.55 if relu else 0.25) < frac < (0.75 if relu else 0.35)  # ~30 % dropped (+ ~50 % by the ReLU)
    assert np.array_equal(sk.asnumpy(mean), sk.asnumpy(mean0)) and np.array_equal(sk.asnumpy(rstd), sk.asnumpy(rstd0))
    r_bwd = 1.0 / keep
    dx, dg, db = F.layernorm_dropout_bwd(adj, x, g, b, mean, rstd, relu, keep, r_bwd, seed)
    adj0 = F.dropout_bwd(adj, keep, r_bwd, seed)
    dx0, dg0, db0, _ = F.layernorm_bwd(adj0, x, g, b, mean0, rstd0, None, 1 if relu else 0)
    assert np.array_equal(sk.asnumpy(dx), sk.asnumpy(dx0))
    assert np.array_equal(sk.asnumpy(dg), sk.asnumpy(dg0)) and np.array_equal(sk.asnumpy(db), sk.asnumpy(db0))


def test_sequential_fuses_layernorm_relu_dropout(sk):
    """nn.Sequential(Linear, LayerNorm, ReLU, Dropout, Linear) under the same seed: the fused
    LayerNorm-ReLU-Dropout kernel and the separate kernels give identical outputs and gradients,
    with one launch less each way."""
    import soket_b200.api as soket
    from soket_b200 import engine as E, nn
    rng = np.random.default_rng(3)
    X = rng.standard_normal((64, 32)).astype("float32")
    W = [rng.standard_normal(s).astype("float32") * 0.3 for s in ((32, 64), (64,), (64,), (64,), (64, 8), (8,))]
    results = []
    try:
        for fuse in (True, False):
            E.set_dropout_fusion(fuse)
            m = nn.Sequential(nn.Linear(32, 64), nn.LayerNorm(64), nn.ReLU(), nn.Dropout(p=0.4), nn.Linear(64, 8))
            for p, w in zip(m.parameters(), W):
                p.data = soket.Tensor(w.copy())
            m.train(True)
            sk.random.seed(77)
            x = soket.Tensor(X, requires_grad=True)
            n0 = sk.launch_count()
            out = m(x)
            n1 = sk.launch_count()
            out.sum().backward()
            results.append((out.numpy(), x.grad.numpy(), [p.grad.numpy() for p in m.parameters()], n1 - n0))
    finally:
        E.set_dropout_fusion(True)
    (o1, g1, p1, l1), (o2, g2, p2, l2) = results
    assert np.array_equal(o1, o2) and np.array_equal(g1, g2)
    for a, c in zip(p1, p2):
        assert np.array_equal(a, c)
    assert l1 == l2 - 1
    assert 0.3 < float((o1 == o2).mean()) and float(np.abs(o1).sum()) > 0
    sk.random.seed(78)                                     # another seed, another mask
    m.train(True)
    assert not np.array_equal(m(soket.Tensor(X)).numpy(), o2)
