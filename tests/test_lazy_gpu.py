"""Lazy mode as a fusion window (reference soket/tensor/tensor.pyx:24-51,
790-810, 1056): with `soket.lazy()` on, float32 elementwise / scalar / unary nodes are recorded and a
whole chain runs as ONE launch (sk_ewise_fused) when a value is needed.  Results must not change: the
case table pinned by the reference (tests/golden/tensor_ops.npz) is replayed with the window open,
and fused chains are compared bit for bit with their op-by-op evaluation."""
import numpy as np
import pytest

from tensor_op_cases import CASES, INT_CASES, make_inputs
from test_tensor_ops_gpu import CRASHES, GOLD, close, seed_of, weights_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_case_table_with_the_window_open(sk, case):
    import soket_b200.api as soket
    name, shapes, fn = case
    if name in CRASHES or f"{name}/backward_error" in GOLD.files:
        pytest.skip("no reference value")
    with soket.lazy():
        xs = [soket.Tensor(a, requires_grad=True) for a in make_inputs(shapes, seed_of(name))]
        out = fn(soket, *xs)
        assert tuple(out.shape) == GOLD[f"{name}/out"].shape          # known before anything has run
        if out.requires_grad:
            w = soket.Tensor(weights_for(GOLD[f"{name}/out"].shape, name))
            (out * w).sum().backward()
        close(out.numpy(), GOLD[f"{name}/out"], f"{name}: forward")
    for i, x in enumerate(xs):
        key = f"{name}/grad{i}"
        if key in GOLD.files:
            assert x.grad is not None, key
            close(x.grad.numpy().astype(GOLD[key].dtype), GOLD[key], key)


@pytest.mark.parametrize("case", INT_CASES, ids=[c[0] for c in INT_CASES])
def test_integer_cases_with_the_window_open(sk, case):
    import soket_b200.api as soket
    name, shapes, fn = case
    with soket.lazy():
        xs = [soket.Tensor(a, requires_grad=True) for a in make_inputs(shapes, seed_of(name))]
        out = fn(soket, *xs)
    close(out.numpy(), GOLD[f"{name}/out"], name)


def _chains(soket):
    relu = soket.nn.functional.relu if hasattr(soket.nn, "functional") and hasattr(soket.nn.functional, "relu") else None
    from soket_b200 import engine as E
    r = E.relu_
    return {
        # Linear's bias add + ReLU + scale + residual-style add, a (cols,) vector operand and scalars
        "bias_relu_residual": lambda a, b, v: r((a * b + v) * 0.5 - a) / (soket.exp(-a) + 1.0),
        # two internal subtrees per binary node -> temporaries
        "two_temps": lambda a, b, v: ((a + b) * (a - b) + (a * b) * (a / (b * b + 1.0))) * v,
        # the LayerNorm tail of forward.pyx:325-352 on precomputed statistics
        "affine": lambda a, b, v: v * ((a - 0.25) * 1.5) + v,
        "pow_forms": lambda a, b, v: (a * a + 1.0) ** 0.5 + (b * b + 0.5) ** -0.5 + (a * a + 2.0) ** 2 - 2.0 ** b,
        "log_exp": lambda a, b, v: soket.log(soket.exp(a) + soket.exp(b)) - 3.0 / (v * v + 1.0),
        "reversed_scalars": lambda a, b, v: 1.0 - (2.0 / (a * a + 1.0)) + (3 - b),
    }


@pytest.mark.parametrize("shape", [(64, 256), (33, 20), (4, 3, 8)])
@pytest.mark.parametrize("chain", ["bias_relu_residual", "two_temps", "affine", "pow_forms", "log_exp", "reversed_scalars"])
def test_fused_chain_is_one_launch_and_bit_identical(sk, chain, shape):
    import soket_b200.api as soket
    from soket_b200 import engine as E
    rng = np.random.default_rng(len(chain) + shape[0])
    an, bn = (rng.standard_normal(shape).astype("float32") for _ in range(2))
    vn = (rng.random(shape[-1]) + 0.5).astype("float32")
    f = _chains(soket)[chain]
    want = f(soket.Tensor(an), soket.Tensor(bn), soket.Tensor(vn)).numpy()          # op by op
    a, b, v = soket.Tensor(an), soket.Tensor(bn), soket.Tensor(vn)
    sk.synchronize()
    E.lazy_stats(reset=True)
    n0 = sk.launch_count()
    with soket.lazy():
        y = f(a, b, v)
        assert sk.launch_count() == n0                         # nothing has run yet
        assert tuple(y.shape) == shape
        got = y.numpy()
    st = E.lazy_stats()
    # a sub-chain of another shape (log_exp: 3 / (v * v + 1) on the (cols,) vector) is a program of its
    # own; everything of the result's shape is one launch
    assert st["programs"] == (2 if chain == "log_exp" else 1) and st["nodes"] >= 4, st
    assert sk.launch_count() - n0 == st["programs"], (sk.launch_count() - n0, st)
    assert np.array_equal(got, want, equal_nan=True), float(np.nanmax(np.abs(got - want)))


def test_long_chain_splits_and_shared_node_is_materialised_once(sk):
    import soket_b200.api as soket
    from soket_b200 import engine as E
    rng = np.random.default_rng(0)
    an = rng.standard_normal((128, 64)).astype("float32")

    def long_chain(a):
        y = a
        for i in range(70):                                    # more steps than one program holds
            y = y * 1.001 + 0.01 if i % 2 else (y - 0.02) / 1.002
        return y
    want = long_chain(soket.Tensor(an)).numpy()
    E.lazy_stats(reset=True)
    with soket.lazy():
        got = long_chain(soket.Tensor(an)).numpy()
    st = E.lazy_stats()
    assert 2 <= st["programs"] <= 6 and st["nodes"] == 140, st
    assert np.array_equal(got, want)
    # a node with two consumers is computed once and read by both
    E.lazy_stats(reset=True)
    with soket.lazy():
        a = soket.Tensor(an)
        t = a * 2.0 + 1.0
        y, z = (t + 1.0) * 3.0, t * t
        yn, zn = y.numpy(), z.numpy()
    st = E.lazy_stats()
    assert st["programs"] == 3 and st["nodes"] == 5, st
    t_ref = an * np.float32(2.0) + np.float32(1.0)
    assert np.array_equal(yn, (t_ref + np.float32(1.0)) * np.float32(3.0)) and np.array_equal(zn, t_ref * t_ref)


def test_backward_through_a_deferred_chain(sk):
    import soket_b200.api as soket
    rng = np.random.default_rng(1)
    an, bn = (rng.standard_normal((32, 48)).astype("float32") for _ in range(2))

    def run(lazy):
        a, b = soket.Tensor(an, requires_grad=True), soket.Tensor(bn, requires_grad=True)
        if lazy:
            with soket.lazy():
                loss = (soket.exp(a * b - 1.0) * (a + 2.0) / (b * b + 1.0)).sum()
                loss.backward()
        else:
            loss = (soket.exp(a * b - 1.0) * (a + 2.0) / (b * b + 1.0)).sum()
            loss.backward()
        return loss.item(), a.grad.numpy(), b.grad.numpy()
    l1, ga1, gb1 = run(True)
    l0, ga0, gb0 = run(False)
    assert l1 == l0
    # the forward values are identical; backward adds its own (also deferred) elementwise chains
    assert np.allclose(ga1, ga0, rtol=1e-6, atol=1e-7) and np.allclose(gb1, gb0, rtol=1e-6, atol=1e-7)
