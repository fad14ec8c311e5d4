"""GPU parity tests of the backend array interface (the intern-table surface,
soket/tensor/ops/intern.pyx:45-76) against NumPy -- the arithmetic layer of the
reference's CPU backend -- with the reference's own call shapes
(forward.pyx / backward.pyx).

Bars (BASELINE.json north_star): bit-exact for indexing / compaction /
broadcast / integer + comparison results and for single IEEE operations
(add, sub, mul, div); <= 1e-5 relative for transcendental and reduced fp32
results.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def rel_err(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if not want.size:
        return 0.0
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    if not np.array_equal(nan_g, nan_w):
        return np.inf
    got, want = np.where(nan_g, 0.0, got), np.where(nan_w, 0.0, want)
    scale = np.maximum(np.abs(want).max(), 1e-30)
    return np.abs(got - want).max() / scale


def assert_close(got, want, rtol=RTOL):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    err = rel_err(got, want)
    assert err <= rtol, f"relative error {err:.3e} > {rtol}"


def assert_exact(got, want):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    if got.dtype.kind == "f":
        assert np.array_equal(got.view(f"u{got.dtype.itemsize}"), want.view(f"u{want.dtype.itemsize}")) or \
            np.array_equal(got, want, equal_nan=True)
    else:
        assert np.array_equal(got, want)


# ----------------------------------------------------------------------------- H2D/D2H
@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "int8", "uint8", "int16",
                                   "uint16", "int32", "uint32", "int64", "uint64", "bool"])
@pytest.mark.parametrize("shape", [(), (1,), (7,), (3, 5), (2, 3, 4), (0,), (4, 0, 2)])
def test_roundtrip(sk, dtype, shape):
    rng = np.random.default_rng(0)
    h = (rng.random(shape) * 100).astype(dtype)
    d = sk.array(h)
    assert d.shape == h.shape and str(d.dtype) == dtype and d.size == h.size
    assert_exact(sk.asnumpy(d), h)


def test_array_from_python_objects(sk):
    assert_exact(sk.asnumpy(sk.array([[1, 2], [3, 4]], "float32")), np.array([[1, 2], [3, 4]], "float32"))
    assert_exact(sk.asnumpy(sk.array(3.5, "float32")), np.array(3.5, "float32"))
    assert sk.array(2, "int32").item() == 2
    assert sk.array(True, "bool").item() is True


# ----------------------------------------------------------------------------- ewise
BIN = [("add", np.add), ("subtract", np.subtract), ("multiply", np.multiply), ("divide", np.divide)]


@pytest.mark.parametrize("name,fn", BIN)
@pytest.mark.parametrize("n", [1, 3, 4, 1000, 4099, 1 << 20])
def test_binary_f32_contiguous_exact(sk, name, fn, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype("float32")
    b = (rng.standard_normal(n).astype("float32") + 2.0).astype("float32")
    got = getattr(sk, name)(sk.array(a), sk.array(b), dtype="float32")
    assert_exact(sk.asnumpy(got), fn(a, b, dtype="float32"))


@pytest.mark.parametrize("name,fn", BIN + [("maximum", np.maximum), ("power", np.power)])
@pytest.mark.parametrize("shapes", [((100, 100), (100,)), ((64, 128), (64, 1)), ((64, 128), (1, 128)),
                                    ((5, 1, 3), (1, 4, 3)), ((8, 12), ()), ((3, 4, 5), (4, 5)),
                                    ((1,), (7, 3)), ((6, 10), (6, 10))])
def test_binary_broadcast(sk, name, fn, shapes):
    rng = np.random.default_rng(1)
    a = (rng.random(shapes[0]) + 0.5).astype("float32")
    b = (rng.random(shapes[1]) + 0.5).astype("float32")
    got = sk.asnumpy(getattr(sk, name)(sk.array(a), sk.array(b)))
    want = fn(a, b)
    if name == "power":
        assert_close(got, want)
    else:
        assert_exact(got, want)


@pytest.mark.parametrize("scalar", [2, 0.5, -0.5, 2.0, 1e-5, 0, -1.0, True])
@pytest.mark.parametrize("name,fn", BIN + [("power", np.power), ("maximum", np.maximum)])
def test_scalar_ops_weak_python_scalars(sk, name, fn, scalar):
    rng = np.random.default_rng(2)
    a = (rng.random((33, 40)) + 0.25).astype("float32")
    if name == "divide" and scalar == 0:
        pytest.skip("division by zero")
    got = sk.asnumpy(getattr(sk, name)(sk.array(a), scalar))
    want = fn(a, scalar)
    if name == "power":
        assert_close(got, want, 2e-6)
    else:
        assert_exact(got, want)
    # reversed operand order (tensor.pyx:1272,1499,1621 put the scalar first)
    got = sk.asnumpy(getattr(sk, name)(scalar, sk.array(a)))
    want = fn(scalar, a)
    if name == "power":
        assert_close(got, want, 2e-6)
    else:
        assert_exact(got, want)


def test_scalar_dtype_promotion(sk):
    ai = np.arange(12, dtype="int32").reshape(3, 4)
    d = sk.array(ai)
    for fn_name, fn in [("add", np.add), ("multiply", np.multiply), ("divide", np.divide), ("subtract", np.subtract)]:
        for s in (2, 1.5):
            got = sk.asnumpy(getattr(sk, fn_name)(d, s))
            want = fn(ai, s)
            assert_exact(got, want)
    # dtype= keyword forces the computation type (forward.pyx:10-13)
    assert_exact(sk.asnumpy(sk.add(d, d, dtype="float32")), np.add(ai, ai, dtype="float32"))
    m = np.array([True, False, True])
    x = np.array([1.5, 2.5, -3.0], "float32")
    assert_exact(sk.asnumpy(sk.multiply(sk.array(m), sk.array(x), dtype="float32")), np.multiply(m, x, dtype="float32"))


@pytest.mark.parametrize("name,fn,tol", [("negative", np.negative, 0), ("exp", np.exp, 2e-6), ("log", np.log, 2e-6)])
@pytest.mark.parametrize("n", [5, 4096, 100003])
def test_unary(sk, name, fn, tol, n):
    rng = np.random.default_rng(3)
    a = (rng.random(n) * 4 + 0.1).astype("float32")
    got = sk.asnumpy(getattr(sk, name)(sk.array(a)))
    want = fn(a)
    if tol == 0:
        assert_exact(got, want)
    else:
        assert_close(got, want, tol)


def test_relu_and_backward(sk):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((100, 100)).astype("float32")
    adj = rng.standard_normal((100, 100)).astype("float32")
    assert_exact(sk.asnumpy(sk.maximum(sk.array(x), 0)), np.maximum(x, 0))
    # reference sequence: greater(x, 0) then multiply(bool, adj, dtype=float32) (backward.pyx:863-867)
    dx, dadj = sk.array(x), sk.array(adj)
    mask = sk.greater(dx, 0)
    assert str(mask.dtype) == "bool"
    got = sk.multiply(mask, dadj, dtype="float32")
    want = np.multiply(np.greater(x, 0), adj, dtype="float32")
    assert_exact(sk.asnumpy(got), want)
    assert_exact(sk.asnumpy(sk.relu_backward(dx, dadj)), want)


@pytest.mark.parametrize("name,fn", [("equal", np.equal), ("not_equal", np.not_equal), ("greater", np.greater),
                                     ("greater_equal", np.greater_equal), ("less", np.less), ("less_equal", np.less_equal)])
def test_compare(sk, name, fn):
    rng = np.random.default_rng(5)
    a = rng.integers(0, 4, (50, 7)).astype("float32")
    b = rng.integers(0, 4, (50, 7)).astype("float32")
    assert_exact(sk.asnumpy(getattr(sk, name)(sk.array(a), sk.array(b))), fn(a, b))
    assert_exact(sk.asnumpy(getattr(sk, name)(sk.array(a), 2)), fn(a, 2))
    # int32 argmax result vs uint8 labels (model.py:69)
    p = rng.integers(0, 10, 64).astype("int32")
    y = rng.integers(0, 10, 64).astype("uint8")
    assert_exact(sk.asnumpy(getattr(sk, name)(sk.array(p), sk.array(y))), fn(p, y))


# ----------------------------------------------------------------------------- views / compaction
def test_views_are_views_and_compaction_is_bit_exact(sk):
    rng = np.random.default_rng(6)
    h = rng.standard_normal((6, 8, 10)).astype("float32")
    d = sk.array(h)
    cases = [
        (sk.transpose(d, (2, 0, 1)), np.transpose(h, (2, 0, 1))),
        (d.T, h.T),
        (d[1:5:2, ::-1, 3], h[1:5:2, ::-1, 3]),
        (d[..., 2], h[..., 2]),
        (d[None, 2], h[None, 2]),
        (sk.reshape(d, (48, 10)), h.reshape(48, 10)),
        (sk.reshape(d.T, (10, 48)), h.T.reshape(10, 48)),
        (sk.reshape(d[:, ::2, :], (6, 40)), h[:, ::2, :].reshape(6, 40)),
        (sk.broadcast_to(d[:, :1, :], (6, 8, 10)), np.broadcast_to(h[:, :1, :], (6, 8, 10))),
        (sk.broadcast_to(sk.array(np.float32(3.0)), (4, 5)), np.broadcast_to(np.float32(3.0), (4, 5))),
        (sk.squeeze(d[:, :1, :], (1,)), np.squeeze(h[:, :1, :], (1,))),
    ]
    for got, want in cases:
        assert got.shape == want.shape
        assert_exact(sk.asnumpy(got), np.ascontiguousarray(want))
        assert_exact(sk.asnumpy(sk.copy(got)), np.ascontiguousarray(want))
    # transposes / slices / broadcasts share storage with the parent
    t = d.T
    assert t.data_ptr == d.data_ptr
    assert sk.broadcast_to(d[0], (3, 8, 10)).strides[0] == 0


@pytest.mark.parametrize("shape", [(1, 1), (31, 33), (128, 64), (1000, 257), (2048, 1024)])
@pytest.mark.parametrize("dtype", ["float32", "float64", "int16", "uint8"])
def test_transpose_compaction_exact(sk, shape, dtype):
    rng = np.random.default_rng(7)
    h = (rng.random(shape) * 255).astype(dtype)
    got = sk.asnumpy(sk.ascontiguousarray(sk.array(h).T))
    assert_exact(got, np.ascontiguousarray(h.T))


@pytest.mark.parametrize("shape", [(4, 4), (64, 64), (68, 132), (100, 784), (784, 100), (1000, 260), (4096, 512)])
def test_transpose64_and_pitched_sources_exact(sk, shape):
    """Multiple-of-4 shapes take the 128-bit swizzled 64 x 64 transpose; also from a pitched
    (column-sliced) parent, and non-multiples fall back to the 32 x 32 kernel -- all bit-exact."""
    rng = np.random.default_rng(shape[0])
    h = rng.standard_normal(shape).astype("float32")
    d = sk.array(h)
    assert_exact(sk.asnumpy(sk.ascontiguousarray(d.T)), np.ascontiguousarray(h.T))
    if shape[1] >= 8:
        assert_exact(sk.asnumpy(sk.ascontiguousarray(d[:, 4:].T)), np.ascontiguousarray(h[:, 4:].T))   # pitched, aligned
        assert_exact(sk.asnumpy(sk.ascontiguousarray(d[:, 1:-3].T)), np.ascontiguousarray(h[:, 1:-3].T))  # misaligned base
        assert_exact(sk.asnumpy(sk.ascontiguousarray(d[:, :-1].T)), np.ascontiguousarray(h[:, :-1].T))    # ragged


@pytest.mark.parametrize("R,C", [(1, 4), (7, 12), (100, 100), (333, 4096), (2048, 8), (5, 10), (64, 1028)])
def test_broadcast_and_pitched_row_compaction_exact(sk, R, C):
    """broadcast_to views (zero strides) and pitched row slices materialise bit-exactly through
    the vectorised row-copy kernel (C % 4 == 0) or the generic gather (otherwise)."""
    rng = np.random.default_rng(R * C)
    row = rng.standard_normal(C).astype("float32")
    col = rng.standard_normal((R, 1)).astype("float32")
    big = rng.standard_normal((R + 3, C + 8)).astype("float32")
    drow, dcol, dbig = sk.array(row), sk.array(col), sk.array(big)
    assert_exact(sk.asnumpy(sk.ascontiguousarray(sk.broadcast_to(drow, (R, C)))), np.broadcast_to(row, (R, C)))
    assert_exact(sk.asnumpy(sk.ascontiguousarray(sk.broadcast_to(dcol, (R, C)))), np.broadcast_to(col, (R, C)))
    assert_exact(sk.asnumpy(sk.ascontiguousarray(sk.broadcast_to(dbig[2:3, 4:4 + C], (R, C)))),
                 np.broadcast_to(big[2:3, 4:4 + C], (R, C)))
    for (r0, c0) in ((0, 0), (1, 4), (2, 3)):
        assert_exact(sk.asnumpy(sk.ascontiguousarray(dbig[r0:r0 + R, c0:c0 + C])), big[r0:r0 + R, c0:c0 + C])
    assert_exact(sk.asnumpy(sk.ascontiguousarray(sk.broadcast_to(dcol[:, 0], (3, R)))), np.broadcast_to(col[:, 0], (3, R)))


def test_astype_matrix(sk):
    rng = np.random.default_rng(8)
    h = (rng.random((17, 9)) * 200 - 50)
    for src in ["float32", "float64", "int32", "uint8", "int64", "bool", "float16"]:
        for dst in ["float32", "float64", "int32", "uint8", "int64", "bool", "float16"]:
            a = h.astype(src)
            if dst.startswith("uint") and src not in ("uint8", "bool"):
                a = np.abs(a)
            assert_exact(sk.asnumpy(sk.array(a).astype(dst)), a.astype(dst))


# ----------------------------------------------------------------------------- indexing
def test_getitem_setitem(sk):
    rng = np.random.default_rng(9)
    h = rng.standard_normal((10, 12)).astype("float32")
    d = sk.array(h)
    assert_exact(sk.asnumpy(d[3]), h[3])
    assert d[3, 4].item() == h[3, 4].item()
    assert_exact(sk.asnumpy(d[2:7, 1::3]), h[2:7, 1::3])
    d[2:4, :] = 7.0
    h[2:4, :] = 7.0
    d[:, 0] = sk.array(np.arange(10, dtype="float32"))
    h[:, 0] = np.arange(10, dtype="float32")
    d[5] = np.ones(12, "float32") * 3
    h[5] = 3
    d[9, 11] = 2
    h[9, 11] = 2
    assert_exact(sk.asnumpy(d), h)


def test_one_hot_via_eye_gather(sk):
    # Device._one_hot: eye(C, None, 0, dtype)[labels], labels uint8 (device.pyx:236-239)
    labels = np.random.default_rng(10).integers(0, 10, 100).astype("uint8")
    got = sk.eye(10, None, 0, "float32")[sk.array(labels)]
    assert_exact(sk.asnumpy(got), np.eye(10, None, 0, "float32")[labels])
    assert_exact(sk.asnumpy(sk.one_hot(sk.array(labels), 10)), np.eye(10, dtype="float32")[labels])
    idx = np.array([[3, 1], [0, -1]], "int64")
    src = np.arange(40, dtype="int32").reshape(4, 5, 2)
    assert_exact(sk.asnumpy(sk.array(src)[sk.array(idx)]), src[idx])


def test_stack(sk):
    rng = np.random.default_rng(11)
    arrs = [rng.standard_normal((3, 4)).astype("float32") for _ in range(5)]
    for axis in (0, 1, 2, -1):
        assert_exact(sk.asnumpy(sk.stack([sk.array(a) for a in arrs], axis=axis)), np.stack(arrs, axis=axis))


def test_creation(sk):
    assert_exact(sk.asnumpy(sk.zeros((3, 4), "float32")), np.zeros((3, 4), "float32"))
    assert_exact(sk.asnumpy(sk.ones((5,), "int32")), np.ones((5,), "int32"))
    assert_exact(sk.asnumpy(sk.full((2, 2), 1.5, "float32")), np.full((2, 2), 1.5, "float32"))
    assert_exact(sk.asnumpy(sk.eye(4, None, 0, "float32")), np.eye(4, None, 0, "float32"))
    assert sk.empty((2, 3), "float16").shape == (2, 3)
    assert_exact(sk.asnumpy(sk.ones((), "float32")), np.ones((), "float32"))


# ----------------------------------------------------------------------------- reductions
RED = [("sum", np.sum), ("mean", np.mean)]


@pytest.mark.parametrize("shape,axes", [((100, 100), (1,)), ((100, 100), (0,)), ((100, 10), (1,)),
                                        ((8192, 512), (1,)), ((8192, 512), (0,)), ((64, 4096), (1,)),
                                        ((3, 5, 7), (0, 2)), ((3, 5, 7), (1,)), ((1 << 20,), None),
                                        ((37, 1031), None), ((5, 1, 6), (1,)), ((4, 300000), (1,)),
                                        ((300000, 4), (0,)), ((7,), (0,))])
@pytest.mark.parametrize("keepdims", [True, False])
def test_sum_mean(sk, shape, axes, keepdims):
    rng = np.random.default_rng(12)
    h = rng.random(shape, dtype=np.float32)
    d = sk.array(h)
    for name, fn in RED:
        got = sk.asnumpy(getattr(sk, name)(d, axes, "float32", None, keepdims))
        want = fn(h, axes, "float32", None, keepdims)
        # NumPy sums strided (non-innermost) axes sequentially in fp32, so its OWN
        # distance to the exact sum grows like sqrt(N)*eps (~1e-5 at N = 8192); the
        # device tree-sum is closer to exact.  Bar: within 2e-6 of NumPy plus
        # NumPy's own distance to the float64 result.
        exact = fn(h.astype(np.float64), axes, None, None, keepdims)
        assert got.shape == want.shape and got.dtype == want.dtype
        assert rel_err(got, want) <= 2e-6 + rel_err(want, exact), (name, rel_err(got, want), rel_err(want, exact))
        assert rel_err(got, exact) <= 2e-6


@pytest.mark.parametrize("shape,axes", [((100, 10), (1,)), ((100, 10), (0,)), ((4096, 1000), (1,)),
                                        ((1 << 18,), None), ((3, 5, 7), (0, 2)), ((6, 5), (-1,))])
def test_max_min_argmax(sk, shape, axes):
    rng = np.random.default_rng(13)
    h = rng.standard_normal(shape).astype("float32")
    d = sk.array(h)
    assert_exact(sk.asnumpy(sk.max(d, axes, None, True)), np.max(h, axes, None, True))
    assert_exact(sk.asnumpy(sk.min(d, axes, None, False)), np.min(h, axes, None, False))
    if axes is None or len(axes) == 1:
        ax = None if axes is None else axes[0]
        for kd in (False, True):
            assert_exact(sk.asnumpy(sk.argmax(d, ax, keepdims=kd)), np.argmax(h, ax, keepdims=kd))
            assert_exact(sk.asnumpy(sk.argmin(d, ax, keepdims=kd)), np.argmin(h, ax, keepdims=kd))


def test_argmax_ties_and_int32_cast(sk):
    h = np.array([[1, 3, 3, 2], [5, 5, 1, 0], [0, 0, 0, 0]], "float32")
    d = sk.array(h)
    got = sk.array(sk.argmax(d, -1, keepdims=False), "int32")  # tensor.pyx:852-859
    assert_exact(sk.asnumpy(got), np.array(np.argmax(h, -1), "int32"))


def test_reduce_dtypes(sk):
    m = np.random.default_rng(14).random((40, 9)) > 0.5
    assert_exact(sk.asnumpy(sk.sum(sk.array(m), (1,), None, None, True)), np.sum(m, (1,), None, None, True))
    assert_close(sk.asnumpy(sk.mean(sk.array(m), None, "float32", None, False)), np.mean(m, None, "float32", None, False))
    i = np.arange(24, dtype="int32").reshape(4, 6)
    assert_exact(sk.asnumpy(sk.sum(sk.array(i), (0,))), np.sum(i, (0,)))
    assert_exact(sk.asnumpy(sk.max(sk.array(i))), np.max(i))


# ----------------------------------------------------------------------------- matmul
def mm_tol(a, b, got, want, rtol=RTOL):
    """|got - want| <= rtol * (|a| @ |b|): the forward-error norm of a dot product."""
    bound = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    err = np.abs(got.astype(np.float64) - want.astype(np.float64))
    worst = (err / np.maximum(bound, 1e-30)).max()
    assert worst <= rtol, f"matmul error {worst:.3e} x (|a|@|b|) > {rtol}"


@pytest.mark.parametrize("M,K,N", [(100, 784, 100), (100, 100, 10), (1, 1, 1), (7, 13, 5), (128, 128, 128),
                                   (257, 129, 65), (100, 100, 100), (512, 300, 384)])
def test_matmul_simt_all_layouts(sk, M, K, N):
    rng = np.random.default_rng(15)
    a = rng.uniform(-1, 1, (M, K)).astype("float32")
    b = rng.uniform(-1, 1, (K, N)).astype("float32")
    adj = rng.uniform(-1, 1, (M, N)).astype("float32")
    da, db, dadj = sk.array(a), sk.array(b), sk.array(adj)
    got = sk.asnumpy(sk.matmul(da, db, dtype="float32", algo=sk.MM_SIMT))
    mm_tol(a, b, got, np.matmul(a, b, dtype="float32"))
    # backward forms (backward.pyx:720-736): adj @ y.T, x.T @ adj with .T VIEWS
    got = sk.asnumpy(sk.matmul(dadj, db.T, algo=sk.MM_SIMT))
    mm_tol(adj, b.T, got, np.matmul(adj, b.T))
    got = sk.asnumpy(sk.matmul(da.T, dadj, algo=sk.MM_SIMT))
    mm_tol(a.T, adj, got, np.matmul(a.T, adj))


def test_matmul_batched_and_vectors(sk):
    rng = np.random.default_rng(16)
    a = rng.standard_normal((3, 2, 5, 7)).astype("float32")
    b = rng.standard_normal((2, 7, 4)).astype("float32")
    got = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b), algo=sk.MM_SIMT))
    assert_close(got, np.matmul(a, b), 1e-5)
    v = rng.standard_normal(7).astype("float32")
    assert_close(sk.asnumpy(sk.matmul(sk.array(a), sk.array(v), algo=sk.MM_SIMT)), np.matmul(a, v), 1e-5)
    assert_close(sk.asnumpy(sk.matmul(sk.array(v), sk.array(b), algo=sk.MM_SIMT)), np.matmul(v, b), 1e-5)


def test_linear_fused_epilogues(sk):
    rng = np.random.default_rng(17)
    x = rng.standard_normal((100, 784)).astype("float32")
    w = (rng.standard_normal((784, 100)) * 0.05).astype("float32")
    b = rng.standard_normal(100).astype("float32")
    pre = np.add(np.matmul(x, w, dtype="float32"), b, dtype="float32")
    got = sk.asnumpy(sk.linear(sk.array(x), sk.array(w), sk.array(b), relu=False, algo=sk.MM_SIMT))
    mm_tol(x, w, got - b, pre - b)
    got = sk.asnumpy(sk.linear(sk.array(x), sk.array(w), sk.array(b), relu=True, algo=sk.MM_SIMT))
    assert np.all((got > 0) == (np.maximum(pre, 0) > 0) | (np.abs(pre) < 1e-4))
    assert_close(got, np.maximum(pre, 0), 1e-5)


# ----------------------------------------------------------------------------- rng
def test_rng_statistics(sk):
    sk.random.seed(123)
    u = sk.asnumpy(sk.random.uniform(-2.0, 3.0, (1000, 1000)))
    assert u.dtype == np.float64 and u.min() >= -2.0 and u.max() < 3.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 25 / 12) < 0.02
    n = sk.asnumpy(sk.random.normal(1.0, 2.0, (1000, 1000)).astype("float32"))
    assert abs(n.mean() - 1.0) < 0.01 and abs(n.std() - 2.0) < 0.01
    b = sk.asnumpy(sk.random.binomial(1, 0.3, (1000, 1000)).astype("float32"))
    assert set(np.unique(b)) <= {0.0, 1.0} and abs(b.mean() - 0.3) < 0.005
    sk.random.seed(123)
    u2 = sk.asnumpy(sk.random.uniform(-2.0, 3.0, (1000, 1000)))
    assert np.array_equal(u, u2)


def test_input_prefetch_on_copy_stream(sk):
    """PinnedBuffer.prefetch_to_device + prefetch_wait: the copy is ordered after the compute work
    queued before it (safe reuse of a staging buffer) and visible to the work queued after the
    wait; two staging buffers, several rounds."""
    n = 1 << 20
    pins = [sk.PinnedBuffer((n,), "float32") for _ in range(2)]
    stage = [sk.empty((n,), "float32") for _ in range(2)]
    rng = np.random.default_rng(9)
    got, want = [], []
    pins[0].array[...] = rng.random(n, dtype=np.float32)
    pins[0].prefetch_to_device(stage[0])
    for k in range(6):
        sk.prefetch_wait()
        want.append(float(pins[k % 2].array.astype(np.float64).sum()))
        nxt = (k + 1) % 2
        sk.synchronize()                       # the host may refill the pinned buffer only after its copy is done
        pins[nxt].array[...] = rng.random(n, dtype=np.float32)
        pins[nxt].prefetch_to_device(stage[nxt])
        heavy = sk.multiply(sk.add(stage[k % 2], 0.0), 1.0)        # compute on the current buffer while the next copies
        got.append(float(sk.asnumpy(sk.sum(heavy, None, "float32", None, False))))
    assert np.allclose(got, want, rtol=1e-5)
