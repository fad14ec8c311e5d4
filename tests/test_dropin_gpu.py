"""Drop-in test: the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/build_ref.py) runs its own Tensor / autodiff / nn / optim code on
`soket.gpu()` with soket_b200 registered in the seam where CuPy sits
(soket_b200.compat), and must agree with its own CPU (NumPy) device:
1e-5 relative on op results, 1e-4 on the loss after a run of training steps.
"""
import numpy as np
import pytest

from oracle import ref_model

pytestmark = pytest.mark.gpu


def rel(got, want):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    return np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)


def test_reference_sees_backend(sk, ref_soket):
    soket = ref_soket
    dev = soket.gpu()
    assert "GPU" in str(dev)
    t = soket.Tensor(np.arange(6, dtype="float32").reshape(2, 3), device=dev)
    assert t.shape == (2, 3)
    assert (t + 1).sum().item() == 21.0


@pytest.mark.parametrize("norm", ["layer", "batch"])
@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_reference_mlpresnet_on_backend_matches_cpu(sk, ref_soket, norm, opt):
    soket = ref_soket
    import soket.nn as nn
    from soket.optim import SGD, Adam
    dim, hidden, nb, C, B, steps = 784, 100, 3, 10, 100, 20
    rng = np.random.default_rng(0)
    from oracle import soket_np
    om = soket_np.MLPResNet(dim, hidden, nb, C, norm=norm)
    for k in om.params:
        if k.endswith(".W"):
            fan = om.params[k].shape[0]
            om.params[k] = (rng.standard_normal(om.params[k].shape) * np.sqrt(2.0 / fan)).astype("float32")
    Xs = rng.random((steps, B, dim), dtype=np.float32)
    ys = rng.integers(0, C, (steps, B)).astype(np.uint8)

    losses = {}
    finals = {}
    for devname in ("cpu", "gpu"):
        dev = soket.cpu() if devname == "cpu" else soket.gpu()
        with dev:
            model = ref_model.build_model(nn, dim, hidden, nb, C, norm=norm, drop_prob=0.0)
            named = ref_model.named_parameters(model, nb)
            for k, t in named.items():
                t.data = soket.Tensor(om.params[k].copy(), device=dev)
            o = SGD(model.parameters(), lr=0.01) if opt == "sgd" else Adam(model.parameters(), lr=0.002)
            crit = nn.SoftmaxCrossEntropyLoss()
            ls = []
            for s in range(steps):
                logits = model(soket.Tensor(Xs[s], device=dev))
                loss = crit(logits, soket.Tensor(ys[s], device=dev))
                loss.backward()
                o.step()
                ls.append(loss.item())
            losses[devname] = np.array(ls)
            finals[devname] = {k: soket.Tensor(t, soket.cpu()) for k, t in named.items()}
    assert np.all(np.isfinite(losses["gpu"]))
    assert np.abs(losses["gpu"] - losses["cpu"]).max() <= 1e-4 * max(1.0, np.abs(losses["cpu"]).max())
    # spot-check final parameters through scalar reductions the reference exposes
    for k in ("lin0.W", "out.W", "blk0.lin1.W", "blk2.n2.g"):
        a = (finals["cpu"][k] * finals["cpu"][k]).sum().item()
        b = (finals["gpu"][k] * finals["gpu"][k]).sum().item()
        assert abs(a - b) <= 1e-4 * max(abs(a), 1e-12), k
