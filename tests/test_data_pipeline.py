"""Input pipeline: soket_b200.utils.data against the reference's
soket/utils/data (loader.py, datasets/mnist.py) -- same batch order, same batch contents.

CPU tests cover the host-side order / shard / idx-gz logic; the gpu tests compare the
device-resident gather path with the per-sample path and with the reference's loader."""
import gzip
import struct

import numpy as np
import pytest


def _write_mnist(tmp_path, n=37, h=28, w=28, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    lab = rng.integers(0, 10, n, dtype=np.uint8)
    fi, fl = str(tmp_path / "img.gz"), str(tmp_path / "lab.gz")
    with gzip.open(fi, "wb") as f:
        f.write(struct.pack(">iiii", 2051, n, h, w) + img.tobytes())
    with gzip.open(fl, "wb") as f:
        f.write(struct.pack(">ii", 2049, n) + lab.tobytes())
    return fi, fl, img, lab


def _ref_to_numpy(soket, t):
    """The reference Tensor has no array accessor: write it into a NumPy buffer through
    Tensor.from_numpy (no copy, tensor.pyx:1073-1084) + __setitem__ (tensor.pyx:948)."""
    if len(t.shape) == 0:
        return np.array(t.item(), dtype=str(t.dtype))
    buf = np.zeros(t.shape, dtype=str(t.dtype))
    view = soket.Tensor.from_numpy(buf)
    view[tuple(slice(None) for _ in t.shape)] = t
    return buf


def _reference_batches(soket, fi, fl, bs, shuffle, seed):
    from soket.transforms import ToTensor as RefToTensor
    from soket.utils.data import DataLoader as RefLoader
    from soket.utils.data.datasets.mnist import MNIST as RefMNIST
    np.random.seed(seed)
    loader = RefLoader(RefMNIST(fi, fl, transforms=RefToTensor(), target_transforms=RefToTensor()),
                       batch_size=bs, shuffle=shuffle)
    out = [(_ref_to_numpy(soket, x), _ref_to_numpy(soket, y)) for x, y in loader]
    return out, loader.ordering


# ------------------------------------------------------------------ CPU: host logic
def test_reference_batches_are_row_gathers_in_ordering(tmp_path, ref_soket):
    """What the resident path relies on: a reference batch == dataset arrays indexed by that
    batch's `ordering` entry (dtype and shape included: (b, 784) float32, (b,) uint8)."""
    from soket_b200.utils.data import MNIST
    fi, fl, _, _ = _write_mnist(tmp_path, n=64, seed=5)
    ds = MNIST(fi, fl)
    for shuffle in (False, True):
        batches, ordering = _reference_batches(ref_soket, fi, fl, 10, shuffle, 3)
        assert len(batches) == 7
        for (x, y), order in zip(batches, ordering):
            assert x.dtype == np.float32 and y.dtype == np.uint8
            assert x.shape == (len(order), 784) and y.shape == (len(order),)
            assert x.tobytes() == ds.data[order].tobytes() and y.tobytes() == ds.targets[order].tobytes()


def test_batch_bounds_equal_numpy_array_split():
    from soket_b200.utils.data import batch_bounds
    for n, m in ((10, 3), (60000, 600), (7, 7), (5, 1), (101, 4), (3, 5), (0, 0)):
        b = batch_bounds(n, m)
        if m == 0:
            assert b == [0]
            continue
        want = [len(c) for c in np.array_split(np.arange(n), m)]
        assert [b[i + 1] - b[i] for i in range(m)] == want and b[0] == 0 and b[-1] == n


def test_shard_bounds_partition_every_batch():
    from soket_b200.utils.data import shard_bounds
    for blen in (100, 8192, 7, 1):
        for world in (1, 2, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_bounds(blen, r, world)
                cover += list(range(lo, hi))
            assert cover == list(range(blen))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_idx_reader_matches_reference_rules(tmp_path):
    from soket_b200.utils.data import MNIST, read_idx_images
    fi, fl, img, lab = _write_mnist(tmp_path)
    ds = MNIST(fi, fl)
    want = img.reshape(len(img), -1).astype(np.float32)
    want /= 255.
    assert ds.data.dtype == np.float32 and ds.data.tobytes() == want.tobytes()
    assert ds.targets.dtype == np.uint8 and np.array_equal(ds.targets, lab)
    assert len(ds) == len(img)
    with pytest.raises(AssertionError):
        read_idx_images(fl)              # label file: wrong magic


def test_idx_reader_bit_equal_to_the_built_reference(tmp_path, ref_soket):
    from soket.utils.data.datasets.mnist import MNIST as RefMNIST
    from soket_b200.utils.data import MNIST
    fi, fl, _, _ = _write_mnist(tmp_path, n=53, seed=3)
    a, b = MNIST(fi, fl), RefMNIST(fi, fl)
    assert a.data.tobytes() == b.data.tobytes() and a.targets.tobytes() == b.targets.tobytes()


def test_loader_order_equals_reference_loader(ref_soket):
    """Same numpy.random stream -> the same `ordering` lists, epoch after epoch."""
    from soket.utils.data import DataLoader as RefLoader
    from soket_b200.utils.data import ArrayDataset, DataLoader
    ds = ArrayDataset(np.zeros((103, 4), np.float32), np.zeros(103, np.uint8))
    for shuffle in (False, True):
        for bs in (1, 10, 50, 103, 200):
            np.random.seed(7)
            ref = RefLoader(ds, batch_size=bs, shuffle=shuffle)
            ref_orders = []
            for _ in range(2):
                iter(ref)
                ref_orders.append([o.copy() for o in ref.ordering])
            np.random.seed(7)
            mine = DataLoader(ds, batch_size=bs, shuffle=shuffle, resident=False)
            assert mine.max_iter == ref.max_iter
            for ep in range(2):
                iter(mine)
                assert len(mine.ordering) == len(ref_orders[ep])
                for x, y in zip(mine.ordering, ref_orders[ep]):
                    assert np.array_equal(x, y)


def test_loader_sharded_ranges_cover_each_batch():
    from soket_b200.utils.data import ArrayDataset, DataLoader
    ds = ArrayDataset(np.arange(50, dtype=np.float32).reshape(25, 2), np.arange(25, dtype=np.uint8))
    full = DataLoader(ds, batch_size=8, resident=False)
    parts = [DataLoader(ds, batch_size=8, resident=False, shard=(r, 3)) for r in range(3)]
    for i in range(full.max_iter):
        lo, hi = full._batch_range(i)
        got = []
        for p in parts:
            a, b = p._batch_range(i)
            got += list(range(a, b))
        assert got == list(range(lo, hi))
    with pytest.raises(ValueError):
        DataLoader(ds, batch_size=8, shard=(3, 3))


def test_resident_detection():
    from soket_b200.transforms import ToTensor, Transform
    from soket_b200.utils.data import ArrayDataset, DataLoader, is_array_backed

    class Flip(Transform):
        def transform(self, x):
            return x[::-1]
    x, y = np.zeros((6, 3), np.float32), np.zeros(6, np.uint8)
    assert is_array_backed(ArrayDataset(x, y))
    assert is_array_backed(ArrayDataset(x, y, ToTensor(), ToTensor()))
    assert not is_array_backed(ArrayDataset(x, y, Flip()))
    with pytest.raises(TypeError):
        DataLoader(ArrayDataset(x, y, Flip()), batch_size=2, resident=True)
    with pytest.raises(ValueError):
        ToTensor()()


# ------------------------------------------------------------------ GPU: batch contents
@pytest.mark.gpu
@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("bs", [1, 16, 37, 100])
def test_resident_batches_bit_equal_per_sample_path(sk, tmp_path, shuffle, bs):
    from soket_b200.transforms import ToTensor
    from soket_b200.utils.data import MNIST, DataLoader
    fi, fl, _, _ = _write_mnist(tmp_path, n=37)
    ds = MNIST(fi, fl, transforms=ToTensor(), target_transforms=ToTensor())
    np.random.seed(11)
    fast = DataLoader(ds, batch_size=bs, shuffle=shuffle)
    assert fast.resident
    got = [(x.numpy(), y.numpy()) for x, y in fast]
    np.random.seed(11)
    slow = DataLoader(ds, batch_size=bs, shuffle=shuffle, resident=False)
    want = [(x.numpy(), y.numpy()) for x, y in slow]
    assert len(got) == len(want) == fast.max_iter
    for (gx, gy), (wx, wy), order in zip(got, want, fast.ordering):
        assert gx.dtype == wx.dtype == np.float32 and gy.dtype == wy.dtype == np.uint8
        assert gx.shape == wx.shape and gy.shape == wy.shape
        assert gx.tobytes() == wx.tobytes() and gy.tobytes() == wy.tobytes()
        assert gx.tobytes() == ds.data[order].tobytes() and gy.tobytes() == ds.targets[order].tobytes()
    n0 = sk.launch_count()
    for _ in fast:
        pass
    assert sk.launch_count() - n0 == 2 * fast.max_iter          # one gather per array per batch


@pytest.mark.gpu
def test_resident_batches_equal_reference_loader(sk, tmp_path, ref_soket):
    """The reference's own loader on its CPU device, same numpy.random seed."""
    from soket_b200.utils.data import MNIST, DataLoader
    fi, fl, _, _ = _write_mnist(tmp_path, n=64, seed=5)
    ref, _ = _reference_batches(ref_soket, fi, fl, 10, True, 3)
    np.random.seed(3)
    got = [(x.numpy(), y.numpy()) for x, y in DataLoader(MNIST(fi, fl), batch_size=10, shuffle=True)]
    assert len(ref) == len(got) == 7
    for (rx, ry), (gx, gy) in zip(ref, got):
        assert rx.shape == gx.shape and ry.shape == gy.shape and rx.dtype == gx.dtype and ry.dtype == gy.dtype
        assert rx.tobytes() == gx.tobytes() and ry.tobytes() == gy.tobytes()


@pytest.mark.gpu
def test_sharded_resident_loader_and_unlabelled_dataset(sk):
    from soket_b200.utils.data import ArrayDataset, DataLoader, ResidentDataset
    rng = np.random.default_rng(0)
    X = rng.random((45, 6), dtype=np.float32)
    y = rng.integers(0, 10, 45).astype(np.uint8)
    np.random.seed(1)
    whole = [(a.numpy(), b.numpy()) for a, b in DataLoader(ArrayDataset(X, y), batch_size=16, shuffle=True)]
    for world in (2, 3):
        pieces = []
        for r in range(world):
            np.random.seed(1)
            pieces.append([(a.numpy(), b.numpy()) for a, b in
                           DataLoader(ArrayDataset(X, y), batch_size=16, shuffle=True, shard=(r, world))])
        for i, (wx, wy) in enumerate(whole):
            assert np.concatenate([p[i][0] for p in pieces]).tobytes() == wx.tobytes()
            assert np.concatenate([p[i][1] for p in pieces]).tobytes() == wy.tobytes()
    only_x = [b.numpy() for b in DataLoader(ArrayDataset(X), batch_size=20)]
    assert np.concatenate(only_x).tobytes() == X.tobytes()
    rd = ResidentDataset(X, y)
    xi, yi = rd[7]
    assert xi.numpy().tobytes() == X[7].tobytes() and int(yi.item()) == int(y[7]) and len(rd) == 45
    gx, gy = rd.gather(np.array([44, 0, 44, -1]))
    assert gx.numpy().tobytes() == X[[44, 0, 44, -1]].tobytes() and np.array_equal(gy.numpy(), y[[44, 0, 44, -1]])


@pytest.mark.gpu
def test_epoch_loop_through_the_loader_matches_oracle(sk):
    """examples/mlp_resnet/model.py:72-97 (mlp_resnet_epoch): a short epoch fed by the resident
    loader gives the oracle's per-step losses on the same batches."""
    import soket_b200.api as soket
    from oracle import ref_model, soket_np as O
    from soket_b200 import nn
    from soket_b200.optim import SGD
    from soket_b200.utils.data import ArrayDataset, DataLoader
    rng = np.random.default_rng(0)
    X = rng.random((300, 784), dtype=np.float32)
    y = rng.integers(0, 10, 300).astype(np.uint8)
    om = O.MLPResNet(784, 64, 2, 10, norm="layer")
    om.init_kaiming(0)
    model = ref_model.build_model(nn, 784, 64, 2, 10, norm="layer", drop_prob=0.0)
    for k, t in ref_model.named_parameters(model, 2).items():
        t.data = soket.Tensor(om.params[k].copy())
    opt, oo = SGD(model.parameters(), lr=0.01), O.SGD(len(om.names()), lr=0.01)
    crit = nn.SoftmaxCrossEntropyLoss()
    np.random.seed(5)
    loader = DataLoader(ArrayDataset(X, y), batch_size=100, shuffle=True)
    batches = iter(loader)                       # shuffling loaders draw `ordering` here (loader.py:56-60)
    for i in range(loader.max_iter):             # (zip() would call iter() again and reshuffle)
        xb, yb = next(batches)
        order = loader.ordering[i]
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        want, _ = om.train_step(X[order], y[order], oo)
        assert abs(loss.item() - want) <= 1e-4 * max(1.0, abs(want))   # north_star: 1e-4 on the loss
