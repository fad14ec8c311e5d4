"""CUDA-graph replay of a whole static-shape training step must equal the
eager engine step for step: same kernels, same order, same buffers -> identical losses."""
import numpy as np
import pytest

from oracle import ref_model

pytestmark = pytest.mark.gpu

DIM, HID, NB, C, B = 784, 64, 2, 10, 100


def build(sk, drop_p, lr, seed=0):
    import soket_b200.api as soket
    from soket_b200 import nn
    from soket_b200.optim import SGD
    rng = np.random.default_rng(seed)
    model = ref_model.build_model(nn, DIM, HID, NB, C, norm="layer", drop_prob=drop_p)
    for t in model.parameters():
        shape = t.shape
        w = (rng.standard_normal(shape) * (np.sqrt(2.0 / shape[0]) if len(shape) == 2 else 0.1)).astype("float32")
        t.data = soket.Tensor(w)
    model.train(True)
    return model, SGD(model.parameters(), lr=lr), nn.SoftmaxCrossEntropyLoss()


def test_graph_replay_equals_eager_steps(sk):
    import soket_b200.api as soket
    from soket_b200.graph import StaticStep
    rng = np.random.default_rng(1)
    Xs = rng.random((8, B, DIM), dtype=np.float32)
    ys = rng.integers(0, C, (8, B)).astype(np.uint8)

    # eager twin: 3 steps on batch 0 (what StaticStep's constructor runs), then batches 1..7
    model, opt, crit = build(sk, 0.0, 0.05)
    want = []
    for i in [0, 0, 0] + list(range(1, 8)):
        loss = crit(model(soket.Tensor(Xs[i])), soket.Tensor(ys[i]))
        loss.backward()
        opt.step()
        want.append(loss.item())

    model, opt, crit = build(sk, 0.0, 0.05)
    xb, yb = soket.Tensor(Xs[0]), soket.Tensor(ys[0])

    def step():
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss
    n0 = sk.launch_count()
    g = StaticStep(step)
    got = [None, None, g.loss.item()]
    per_step = (sk.launch_count() - n0) // 3
    for i in range(1, 8):
        xb._data[:] = sk.array(Xs[i])
        yb._data[:] = sk.array(ys[i])
        n1 = sk.launch_count()
        g.launch()
        assert sk.launch_count() - n1 <= 1           # the whole step is one graph launch
        got.append(g.loss.item())
    assert per_step > 20
    assert got[2:] == want[2:], (got, want)          # bit-identical: same kernels on the same data
    g.close()


def test_graph_replays_draw_fresh_dropout_masks(sk):
    import soket_b200.api as soket
    from soket_b200.graph import StaticStep
    rng = np.random.default_rng(2)
    X = rng.random((B, DIM), dtype=np.float32)
    y = rng.integers(0, C, B).astype(np.uint8)
    model, opt, crit = build(sk, 0.5, 0.0)           # lr = 0: only the masks change between replays
    xb, yb = soket.Tensor(X), soket.Tensor(y)

    def step():
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss
    g = StaticStep(step)
    losses = []
    for _ in range(6):
        g.launch()
        losses.append(g.loss.item())
    assert len(set(losses)) == 6, losses
    g.close()


def test_graph_capture_refuses_adam(sk):
    import soket_b200.api as soket
    from soket_b200 import nn
    from soket_b200.graph import StaticStep
    from soket_b200.optim import Adam
    model, _, crit = build(sk, 0.0, 0.0)
    opt = Adam(model.parameters(), lr=1e-3)
    rng = np.random.default_rng(3)
    xb = soket.Tensor(rng.random((B, DIM), dtype=np.float32))
    yb = soket.Tensor(rng.integers(0, C, B).astype(np.uint8))

    def step():
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss
    with pytest.raises(RuntimeError, match="Adam"):
        StaticStep(step)
    # the engine is still usable afterwards
    assert np.isfinite(step().item())


def test_capturable_adam_graph_replay_equals_eager_default_adam(sk):
    """Adam(capturable=True) keeps beta^t on the device (sk_adam_step_dev / sk_adam_bias_advance):
    a replayed step must equal the DEFAULT Adam's eager step bit for bit -- losses through ten
    steps and every parameter at the end -- weight decay included."""
    import soket_b200.api as soket
    from soket_b200.graph import StaticStep
    from soket_b200.optim import Adam
    rng = np.random.default_rng(4)
    Xs = rng.random((8, B, DIM), dtype=np.float32)
    ys = rng.integers(0, C, (8, B)).astype(np.uint8)
    order = [0, 0, 0] + list(range(1, 8))

    model, _, crit = build(sk, 0.0, 0.0)
    opt = Adam(model.parameters(), lr=1e-3, weight_decay=0.01)
    want = []
    for i in order:
        loss = crit(model(soket.Tensor(Xs[i])), soket.Tensor(ys[i]))
        loss.backward()
        opt.step()
        want.append(loss.item())
    want_params = [p.numpy() for p in model.parameters()]

    model, _, crit = build(sk, 0.0, 0.0)
    opt = Adam(model.parameters(), lr=1e-3, weight_decay=0.01, capturable=True)
    xb, yb = soket.Tensor(Xs[0]), soket.Tensor(ys[0])

    def step():
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss
    g = StaticStep(step)
    got = [None, None, g.loss.item()]
    for i in range(1, 8):
        xb._data[:] = sk.array(Xs[i])
        yb._data[:] = sk.array(ys[i])
        g.launch()
        got.append(g.loss.item())
    assert got[2:] == want[2:], (got, want)
    for p, w in zip(model.parameters(), want_params):
        assert p.numpy().tobytes() == w.tobytes()
    # the device-side products are beta^(t+1) after t steps, as the host recurrence gives them
    b1, b2 = 0.9, 0.999
    h1, h2 = b1, b2
    for _ in order:
        h1 *= b1
        h2 *= b2
    assert sk.asnumpy(opt._bias_dev).tolist() == [h1, h2]
    g.close()


def test_capturable_adam_eager_equals_default_adam(sk):
    """No graph involved: the two forms of the kernel argument give identical updates."""
    import soket_b200.api as soket
    from soket_b200.optim import Adam
    rng = np.random.default_rng(5)
    shapes = [(33, 17), (17,), (1000, 3)]
    init = [rng.standard_normal(s).astype("float32") for s in shapes]
    grads = [[rng.standard_normal(s).astype("float32") for s in shapes] for _ in range(4)]
    finals = []
    for capturable in (False, True):
        ps = [soket.Tensor(w.copy(), requires_grad=True) for w in init]
        opt = Adam(ps, lr=1e-2, capturable=capturable)
        for step_grads in grads:
            loss = None
            for p, gr in zip(ps, step_grads):
                term = (p * soket.Tensor(gr)).sum()
                loss = term if loss is None else loss + term
            loss.backward()
            opt.step()
        finals.append([p.numpy() for p in ps])
    for a, b in zip(*finals):
        assert a.tobytes() == b.tobytes()
