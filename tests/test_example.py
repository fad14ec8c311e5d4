"""examples/mlp_resnet/train_mnist.py -- the counterpart of the reference's
examples/mlp_resnet/model.py -- against the oracle: accuracy (model.py:61-69), the epoch
loop (model.py:72-97) in training and evaluation, and the command-line driver."""
import importlib.util
import os

import numpy as np
import pytest

from test_data_pipeline import _write_mnist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _example():
    spec = importlib.util.spec_from_file_location(
        "soket_b200_example_train_mnist", os.path.join(ROOT, "examples", "mlp_resnet", "train_mnist.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _logits_with_ties(seed=0, rows=257, classes=10):
    rng = np.random.default_rng(seed)
    Z = (4 * rng.standard_normal((rows, classes))).astype(np.float32)
    Z[3] = 0.0                      # all equal: first index wins
    Z[5, 2] = Z[5, 7] = Z[5].max() + 1.0   # two-way tie for the maximum
    y = rng.integers(0, classes, rows).astype(np.uint8)
    y[3], y[5] = 0, 2
    y[: rows // 2] = Z[: rows // 2].argmax(-1).astype(np.uint8)   # about half right
    return Z, y


def test_oracle_accuracy_equals_reference_expression(ref_soket):
    """The NumPy restatement against the reference's own Tensor ops on its CPU device."""
    from oracle import soket_np as O
    soket = ref_soket
    Z, y = _logits_with_ties()
    Zt, yt = soket.Tensor(Z), soket.Tensor(y)
    e = soket.exp(Zt - Zt.max(-1, keepdims=True))
    sm = e / e.sum(-1, keepdims=True)
    want = (sm.argmax(-1) == yt).mean(dtype=soket.float32).item()
    assert O.accuracy(Z, y) == want
    assert 0.45 < want < 0.75


@pytest.mark.gpu
def test_accuracy_matches_oracle(sk):
    import soket_b200.api as soket
    from oracle import soket_np as O
    ex = _example()
    for seed, rows in ((0, 257), (1, 100), (2, 8192)):
        Z, y = _logits_with_ties(seed, rows)
        got = ex.mlp_resnet_get_accuracy(soket.Tensor(Z), soket.Tensor(y))
        assert isinstance(got, float)
        assert got == O.accuracy(Z, y)          # a count / rows: exact


@pytest.mark.gpu
def test_epoch_loop_train_and_eval_match_oracle(sk, tmp_path):
    """mlp_resnet_epoch over the MNIST reader + resident loader: per-epoch mean loss within
    1e-4 (north_star) and mean accuracy equal to the oracle's on the same batches."""
    import soket_b200.api as soket
    from oracle import ref_model, soket_np as O
    from soket_b200.optim import SGD
    from soket_b200.utils.data import MNIST, DataLoader
    ex = _example()
    fi, fl, _, _ = _write_mnist(tmp_path, n=300, seed=2)
    ds = MNIST(fi, fl)
    om = O.MLPResNet(784, 64, 2, 10, norm="layer")
    om.init_kaiming(0)
    rng = np.random.default_rng(0)
    for k in om.linear_weight_names():          # He-scale weights: non-trivial logits
        shp = om.params[k].shape
        om.params[k] = (rng.standard_normal(shp) * np.sqrt(2.0 / shp[0])).astype("float32")
    model = ex.MLPResNet(784, hidden_dim=64, num_blocks=2, num_classes=10, drop_prob=0.0, train_inner=True)
    for k, t in ref_model.named_parameters(model, 2).items():
        t.data = soket.Tensor(om.params[k].copy())
    opt, oo = SGD(model.parameters(), lr=0.05), O.SGD(len(om.names()), lr=0.05)
    model.train(True)
    np.random.seed(9)
    loader = DataLoader(ds, batch_size=100, shuffle=True)
    loss, acc = ex.mlp_resnet_epoch(model, loader, opt)
    want_loss = want_acc = 0.0
    for order in loader.ordering:
        step_loss, logits = om.train_step(ds.data[order], ds.targets[order], oo)
        want_loss += step_loss
        want_acc += O.accuracy(logits, ds.targets[order])
    want_loss /= 3
    want_acc /= 3
    assert abs(loss - want_loss) <= 1e-4 * max(1.0, abs(want_loss))
    assert abs(acc - want_acc) <= 1.0 / 100 + 1e-9      # at most one near-tie flip per 300 rows
    # evaluation pass (no optimiser), sequential batches
    model.train(False)
    test_loader = DataLoader(ds, batch_size=150, shuffle=False)
    loss, acc = ex.mlp_resnet_epoch(model, test_loader)
    want_loss = want_acc = 0.0
    for order in test_loader.ordering:
        logits = om.forward(ds.data[order])
        want_loss += float(om.loss(logits, ds.targets[order]))
        want_acc += O.accuracy(logits, ds.targets[order])
    assert abs(loss - want_loss / 2) <= 1e-4 * max(1.0, abs(want_loss / 2))
    assert abs(acc - want_acc / 2) <= 1.0 / 150 + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("extra,min_acc", [
    ([], 0.0), (["--graph"], 0.0),
    (["--optimizer", "adam", "--lr", "0.001", "--weight-decay", "0.001"], 0.8),
    (["--optimizer", "adam", "--lr", "0.001", "--weight-decay", "0.001", "--graph"], 0.8)])
def test_driver_on_synthetic_mnist(sk, extra, min_acc, capsys):
    """The command-line driver end to end (resident loader, train epochs, eval epoch).  With the
    reference's variance-as-std kaiming init (quirk Q9) plain SGD barely moves in 60 steps (the
    oracle reaches 0.2-0.26 test accuracy), Adam reaches 1.0 on this separable data."""
    ex = _example()
    argv = ["--synthetic", "3000", "--batch-size", "100", "--epochs", "2", "--optimizer", "sgd", "--lr", "0.01",
            "--weight-decay", "0.0", "--train-inner", "--hidden-dim", "64"] + extra
    loss, acc = ex.main(argv)
    out = capsys.readouterr().out
    assert "Epoch: 1, train loss:" in out and "Test loss:" in out
    assert np.isfinite(loss) and min_acc <= acc <= 1.0
