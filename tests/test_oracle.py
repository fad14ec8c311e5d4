"""CPU tests: the NumPy restatement (oracle/soket_np.py) is pinned to the reference.

  * against the committed golden vectors (tests/golden/*.npz, produced by the BUILT
    reference through tests/golden/make_golden.py) -- always runs;
  * against the built reference itself (oracle/_ref) when it is present -- bit-exact
    loss trajectories for LayerNorm / BatchNorm x SGD / Adam.
"""
import os

import numpy as np
import pytest

from oracle import soket_np as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.mark.parametrize("norm", ["layer", "batch"])
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_oracle_reproduces_reference_training_bit_exact(norm, opt):
    g = load(f"mlpresnet_{norm}_{opt}.npz")
    X, y = g["X"], g["y"]
    steps, B, dim = X.shape
    hidden = g["init/lin0.W"].shape[1]
    C = g["init/out.W"].shape[1]
    nb = sum(1 for k in g if k.startswith("init/blk") and k.endswith("lin1.W"))
    om = O.MLPResNet(dim, hidden, nb, C, norm=norm)
    for k in om.names():
        om.params[k] = g[f"init/{k}"].copy()
    lr, wd = g["hyper"]
    names = om.names()
    oo = O.SGD(len(names), lr=float(lr), weight_decay=float(wd)) if opt == "sgd" else \
        O.Adam(len(names), lr=float(lr), weight_decay=float(wd))
    losses = []
    for s in range(steps):
        if s == 0:
            logits = om.forward(X[0])
            om.loss(logits, y[0])
            G = om.backward()
            assert np.array_equal(logits, g["logits0"])
            for k in names:
                assert np.array_equal(np.asarray(G[k]).reshape(g[f"grad0/{k}"].shape), g[f"grad0/{k}"]), k
        l, _ = om.train_step(X[s], y[s], oo)
        losses.append(l)
    assert np.array_equal(np.array(losses), g["losses"])
    for k in names:
        assert np.array_equal(om.params[k].reshape(g[f"final/{k}"].shape), g[f"final/{k}"]), k


def test_oracle_ops_match_reference_vectors():
    g = load("ops.npz")
    x, gam, bet, coef = g["ln_x"], g["ln_g"], g["ln_b"], g["ln_coef"]
    y, xs, rv, norm, _, _ = O.norm_fwd(x, gam, bet, (1,), 1e-5, True)
    assert np.array_equal(y, g["ln_y"])
    dz, dg, db = O.norm_bwd(coef, gam, xs, rv, norm, (1,), x.shape[1], True)
    assert np.array_equal(dz, g["ln_dx"]) and np.array_equal(dg, g["ln_dg"]) and np.array_equal(db, g["ln_db"])
    x, gam, bet, coef = g["bn_x"], g["bn_g"], g["bn_b"], g["bn_coef"]
    y, xs, rv, norm, _, _ = O.norm_fwd(x, gam, bet, (0,), 1e-5, False, np.array(0.0, "float32"), np.array(1.0, "float32"), 0.1)
    assert np.array_equal(y, g["bn_y"])
    dz, dg, db = O.norm_bwd(coef, gam, xs, rv, norm, (0,), x.shape[0], False)
    assert np.array_equal(dz, g["bn_dx"]) and np.array_equal(dg, g["bn_dg"]) and np.array_equal(db, g["bn_db"])
    oh = O.one_hot(g["ce_labels"], 10)
    assert np.array_equal(O.sxent_fwd(g["ce_logits"], oh), g["ce_loss"])
    assert np.array_equal(O.sxent_bwd(np.ones((), "float32"), g["ce_logits"], oh), g["ce_dlogits"])
    a, w = g["a"], g["w"]
    mm = O.linear_fwd(a, w)
    assert np.array_equal(mm, g["matmul"])
    adj = np.multiply(mm, 1.0) + np.multiply(mm, 1.0)  # d/dmm of sum(mm*mm): two partials mm, mm
    da, dw = O.matmul_bwd(adj.astype("float32"), a, w)
    assert np.allclose(da, g["matmul_da"], rtol=1e-6) and np.allclose(dw, g["matmul_dw"], rtol=1e-6)
    assert np.allclose(g["sgd_w2"], [0.512, 1.024], rtol=1e-6)   # known answer


@pytest.mark.parametrize("norm", ["layer", "batch"])
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_oracle_matches_built_reference_live(ref_soket, norm, opt):
    """Fresh seeds, larger model than the fixtures; needs oracle/_ref."""
    from oracle import ref_model
    soket = ref_soket
    import soket.nn as nn
    from soket.optim import SGD, Adam
    dim, hidden, nb, C, B = 40, 32, 3, 10, 16
    rng = np.random.default_rng(7)
    om = O.MLPResNet(dim, hidden, nb, C, norm=norm)
    for k in om.params:
        if k.endswith(".W"):
            om.params[k] = (rng.standard_normal(om.params[k].shape) * 0.3).astype("float32")
        else:
            om.params[k] = (om.params[k] + 0.1 * rng.standard_normal(om.params[k].shape)).astype("float32")
    model = ref_model.build_model(nn, dim, hidden, nb, C, norm=norm, drop_prob=0.0)
    named = ref_model.named_parameters(model, nb)
    for k, t in named.items():
        t.data = soket.Tensor(om.params[k].copy())
    names = om.names()
    if opt == "sgd":
        ro, oo = SGD(model.parameters(), lr=0.05), O.SGD(len(names), lr=0.05)
    else:
        ro, oo = Adam(model.parameters(), lr=0.01, weight_decay=0.001), O.Adam(len(names), lr=0.01, weight_decay=0.001)
    crit = nn.SoftmaxCrossEntropyLoss()
    for _ in range(6):
        X = rng.random((B, dim), dtype=np.float32)
        y = rng.integers(0, C, B).astype(np.uint8)
        loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
        loss.backward()
        ro.step()
        l, _ = om.train_step(X, y, oo)
        assert loss.item() == l


def test_data_parallel_shards_reproduce_global_gradient():
    """with mean-reduced CE, averaging the gradients of W equal row
    shards reproduces the global-batch gradient (LayerNorm model; BatchNorm uses
    per-shard statistics and is NOT expected to)."""
    dim, hidden, nb, C, B, W = 30, 16, 2, 10, 32, 4
    rng = np.random.default_rng(3)
    om = O.MLPResNet(dim, hidden, nb, C, norm="layer")
    for k in om.params:
        if k.endswith(".W"):
            om.params[k] = (rng.standard_normal(om.params[k].shape) * 0.3).astype("float32")
    X = rng.random((B, dim), dtype=np.float32)
    y = rng.integers(0, C, B).astype(np.uint8)
    om.loss(om.forward(X), y)
    G = {k: np.array(v, dtype=np.float64) for k, v in om.backward().items()}
    acc = {k: 0.0 for k in G}
    from soket_b200.dp import shard_rows
    for r in range(W):
        sl = shard_rows(B, r, W)
        om.loss(om.forward(X[sl]), y[sl])
        for k, v in om.backward().items():
            acc[k] = acc[k] + np.asarray(v, dtype=np.float64) / W
    for k in G:
        assert np.abs(acc[k] - G[k]).max() <= 1e-5 * max(np.abs(G[k]).max(), 1e-6), k
