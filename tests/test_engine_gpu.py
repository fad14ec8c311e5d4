"""GPU parity of the resident / fused training path (soket_b200.engine, .nn, .optim)
against the NumPy oracle (bit-exact restatement of the built reference):
per-step loss, gradients and parameters of the north-star model
(examples/mlp_resnet/model.py:40-58) for LayerNorm / BatchNorm x SGD / Adam, with
fusion on (product path) and off (reference op sequence on the same kernels).

Tolerances (BASELINE.json north_star): 1e-5 relative on fp32 op results
(gradients), 1e-4 on the loss after a run of training steps.
"""
import numpy as np
import pytest

from oracle import ref_model, soket_np as O

pytestmark = pytest.mark.gpu


def rel(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def make_pair(sk, norm, dim, hidden, nb, C, seed=0):
    import soket_b200.api as soket
    from soket_b200 import nn
    rng = np.random.default_rng(seed)
    om = O.MLPResNet(dim, hidden, nb, C, norm=norm)
    for k in om.params:
        if k.endswith(".W"):
            fan = om.params[k].shape[0]
            om.params[k] = (rng.standard_normal(om.params[k].shape) * np.sqrt(2.0 / fan)).astype("float32")
        elif ".n" in k and k.endswith(".g"):
            om.params[k] = (1 + 0.1 * rng.standard_normal(om.params[k].shape)).astype("float32")
        else:
            om.params[k] = (0.1 * rng.standard_normal(om.params[k].shape)).astype("float32")
    model = ref_model.build_model(nn, dim, hidden, nb, C, norm=norm, drop_prob=0.0)
    named = ref_model.named_parameters(model, nb)
    for k, t in named.items():
        t.data = soket.Tensor(om.params[k].copy())
    return om, model, named


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("norm", ["layer", "batch"])
def test_gradients_match_oracle(sk, norm, fuse):
    import soket_b200.api as soket
    from soket_b200 import nn
    nn.set_fusion(fuse)
    try:
        dim, hidden, nb, C, B = 784, 100, 3, 10, 100
        om, model, named = make_pair(sk, norm, dim, hidden, nb, C)
        rng = np.random.default_rng(1)
        X = rng.random((B, dim), dtype=np.float32)
        y = rng.integers(0, C, B).astype(np.uint8)
        logits = model(soket.Tensor(X))
        loss = nn.SoftmaxCrossEntropyLoss()(logits, soket.Tensor(y))
        loss.backward()
        want_logits = om.forward(X)
        want_loss = om.loss(want_logits, y)
        G = om.backward()
        assert rel(logits.numpy(), want_logits) <= 1e-5
        assert abs(loss.item() - float(want_loss)) <= 1e-5 * max(1.0, abs(float(want_loss)))
        # a bias in front of BatchNorm has an exactly-zero gradient in real arithmetic:
        # both sides hold ~1e-8 rounding residue there, hence the absolute floor
        gscale = max(np.abs(g).max() for g in G.values())
        for k, t in named.items():
            assert t.grad is not None, k
            assert t.grad.shape == om.params[k].shape, k
            err = np.abs(t.grad.numpy().astype(np.float64) - G[k]).max()
            assert err <= 2e-5 * np.abs(G[k]).max() + 1e-6 * gscale, (k, err)
    finally:
        nn.set_fusion(True)


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("norm", ["layer", "batch"])
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_training_trajectory_matches_oracle(sk, norm, opt, fuse):
    import soket_b200.api as soket
    from soket_b200 import nn
    from soket_b200.optim import SGD, Adam
    nn.set_fusion(fuse)
    try:
        dim, hidden, nb, C, B, steps = 784, 100, 3, 10, 100, 30
        om, model, named = make_pair(sk, norm, dim, hidden, nb, C)
        names = om.names()
        if opt == "sgd":
            oo = O.SGD(len(names), lr=0.01)
            do = SGD(model.parameters(), lr=0.01)
        else:
            oo = O.Adam(len(names), lr=0.001, weight_decay=0.001)
            do = Adam(model.parameters(), lr=0.001, weight_decay=0.001)
        crit = nn.SoftmaxCrossEntropyLoss()
        rng = np.random.default_rng(2)
        got, want = [], []
        for s in range(steps):
            X = rng.random((B, dim), dtype=np.float32)
            y = rng.integers(0, C, B).astype(np.uint8)
            loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
            loss.backward()
            do.step()
            got.append(loss.item())
            l, _ = om.train_step(X, y, oo)
            want.append(l)
        got, want = np.array(got), np.array(want)
        # The CPU path's OWN sensitivity to fp32 rounding: the same oracle with every
        # matmul replaced by an equally valid fp32 evaluation in a different summation
        # order (float64 result + sqrt(K)*2^-24 random-walk rounding error, see
        # oracle/soket_np.py).  ReLU sign flips and Adam's g/(|g|+eps) normalisation of
        # near-zero gradients amplify ulp-level differences; the device path cannot be
        # closer to NumPy than NumPy is to itself.  Bar: 1e-4 (north_star) + 2x that.
        om2, _, _ = make_pair(sk, norm, dim, hidden, nb, C)
        oo2 = O.SGD(len(names), lr=0.01) if opt == "sgd" else O.Adam(len(names), lr=0.001, weight_decay=0.001)
        rng = np.random.default_rng(2)
        pert = []
        with O.matmul_mode("order"):
            for s in range(steps):
                X = rng.random((B, dim), dtype=np.float32)
                y = rng.integers(0, C, B).astype(np.uint8)
                pert.append(om2.train_step(X, y, oo2)[0])
        sens = np.abs(np.array(pert) - want)
        err = np.abs(got - want)
        # Past the step where NumPy-vs-NumPy itself differs by more than 1e-4 the
        # trajectory is numerically undetermined (BatchNorm + Adam is chaotic at this
        # size); compare over the determined horizon only.
        over = np.nonzero(sens > 1e-4)[0]
        horizon = int(over[0]) if len(over) else steps
        assert horizon >= 5, horizon
        bound = 1e-4 * max(1.0, np.abs(want).max()) + 2 * sens[:horizon].max()
        assert np.all(err[:horizon] <= bound), (horizon, err[:horizon].max(), sens[:horizon].max())
        if horizon < steps:
            return
        for k in ("lin0.W", "blk1.lin2.W", "blk2.n1.g", "out.b"):
            assert rel(named[k].numpy(), om.params[k]) <= 1e-4 + 2 * rel(om2.params[k], om.params[k]), k
    finally:
        nn.set_fusion(True)


def test_module_discovery_quirk_q1(sk):
    """Residual hides its layers from parameters()/modules() (prototypes.pyx:261,268);
    the `self.fn`-retaining subclass exposes them: 4 vs 4 + 8 * blocks tensors."""
    from soket_b200 import nn
    m = ref_model.build_model(nn, 16, 8, 3, 4, retain_fn=False)
    assert len(list(m.parameters())) == 4
    m = ref_model.build_model(nn, 16, 8, 3, 4, retain_fn=True)
    assert len(list(m.parameters())) == 4 + 8 * 3


def test_tensor_ops_and_quirks(sk):
    import soket_b200.api as soket
    a = soket.Tensor(np.arange(6, dtype="float32").reshape(2, 3), requires_grad=True)
    b = soket.Tensor(np.ones((3,), "float32") * 2, requires_grad=True)
    z = ((a * b + 1.0) / 2.0 - a ** 2).sum()
    z.backward()
    x = np.arange(6, dtype="float32").reshape(2, 3)
    assert np.allclose(a.grad.numpy(), 1.0 - 2 * x)
    assert np.allclose(b.grad.numpy(), (x / 2).sum(0))          # broadcast gradient summed back
    assert soket.Tensor(np.zeros((2, 3, 4), "float32")).T.shape == (4, 3, 2)   # quirk Q7
    p = soket.Tensor([1.0, 2.0, 3.0]); q = soket.Tensor([2.0, 2.0, 2.0])
    assert (p <= q).numpy().tolist() == [False, True, True]     # quirk Q6: `<=` is `>=`
    w = soket.Tensor([1.0, 2.0], requires_grad=True)
    from soket_b200.optim import SGD
    o = SGD([w], lr=0.1, momentum=0.9)
    for _ in range(3):
        (w * w).sum().backward()
        o.step()
    assert np.allclose(w.numpy(), [0.512, 1.024], rtol=1e-6)    # quirk Q2: plain SGD (known answer)
    assert (soket.Tensor(np.array(3.0, 'float32')) + 2).item() == 5.0
    m = soket.Tensor(np.arange(12, dtype="float32").reshape(3, 4), requires_grad=True)
    m.mean().backward()
    assert np.allclose(m.grad.numpy(), 1.0)                     # quirk Q5: full mean backward unscaled
    s = soket.Tensor(np.arange(12, dtype="float32").reshape(3, 4), requires_grad=True)
    s[1:, ::2].sum().backward()
    want = np.zeros((3, 4), "float32"); want[1:, ::2] = 1
    assert np.array_equal(s.grad.numpy(), want)
    assert soket.Tensor(np.array([[1, 5, 2]], "float32")).argmax(-1).numpy().tolist() == [1]
