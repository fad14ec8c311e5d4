"""Tensor-level op cases shared by the reference (oracle/_ref, CPU) and soket_b200 (GPU):
each case is (name, input shapes, fn(soket, *tensors) -> Tensor).  `soket` is the namespace
under test (the built reference's `soket` or `soket_b200.api`); the bodies use only the public
Tensor API of soket/tensor/tensor.pyx, so the same code runs on both.  SURVEY.md section 8(a)
"Tensor-level methods on the path" + section 8(f)-3."""
import numpy as np


def _t(soket, a):
    return soket.Tensor(np.asarray(a))


CASES = [
    # ---- elementwise, tensor (x) tensor with broadcasting (tensor.pyx:1101-1583) -----------
    ("add_bcast_row", [(5, 7), (7,)], lambda s, a, b: a + b),
    ("add_bcast_col", [(5, 7), (5, 1)], lambda s, a, b: a + b),
    ("add_bcast_both", [(5, 1, 4), (3, 1)], lambda s, a, b: a + b),
    ("sub", [(4, 6), (4, 6)], lambda s, a, b: a - b),
    ("sub_bcast", [(6,), (4, 6)], lambda s, a, b: a - b),
    ("mul", [(4, 6), (4, 6)], lambda s, a, b: a * b),
    ("mul_bcast_scalar_like", [(4, 6), (1, 1)], lambda s, a, b: a * b),
    ("div", [(4, 6), (4, 6)], lambda s, a, b: a / (b * b + 1.0)),
    ("div_bcast", [(4, 6), (6,)], lambda s, a, b: a / (b * b + 0.5)),
    ("pow_tensor", [(3, 5), (3, 5)], lambda s, a, b: (a * a + 0.5) ** b),
    ("neg", [(3, 5)], lambda s, a: -a),
    # ---- scalar operands, both sides (tensor.pyx:231-237, _rsub/_rdiv/_rpow) ------------
    ("add_scalar", [(3, 5)], lambda s, a: a + 2.5),
    ("radd_scalar", [(3, 5)], lambda s, a: 2.5 + a),
    ("sub_scalar", [(3, 5)], lambda s, a: a - 1.25),
    ("rsub_scalar", [(3, 5)], lambda s, a: 1.25 - a),
    ("mul_scalar_int", [(3, 5)], lambda s, a: a * 3),
    ("rmul_scalar", [(3, 5)], lambda s, a: 0.5 * a),
    ("div_scalar", [(3, 5)], lambda s, a: a / 4.0),
    ("rdiv_scalar", [(3, 5)], lambda s, a: 2.0 / (a * a + 1.0)),
    ("pow_scalar", [(3, 5)], lambda s, a: (a * a + 0.5) ** 1.5),
    ("pow_scalar_int", [(3, 5)], lambda s, a: a ** 3),
    ("rpow_scalar", [(3, 5)], lambda s, a: 2.0 ** a),
    # ---- reductions (tensor.pyx:1680-1822) ----------------------------------------------------
    ("sum_all", [(4, 5, 6)], lambda s, a: a.sum()),
    ("sum_axis0", [(4, 5, 6)], lambda s, a: a.sum(0)),
    ("sum_axis_last_keep", [(4, 5, 6)], lambda s, a: a.sum(-1, keepdims=True)),
    ("sum_axes_02", [(4, 5, 6)], lambda s, a: a.sum(0, 2)),
    ("sum_axes_tuple", [(4, 5, 6)], lambda s, a: a.sum((1, 2))),
    ("mean_axis1", [(4, 5, 6)], lambda s, a: a.mean(1)),
    ("mean_axis0_keep", [(4, 5, 6)], lambda s, a: a.mean(0, keepdims=True)),
    ("mean_all", [(4, 5)], lambda s, a: a.mean()),
    ("max_axis1", [(4, 5, 6)], lambda s, a: a.max(1)),
    ("max_keep", [(4, 5, 6)], lambda s, a: a.max(-1, keepdims=True)),
    ("max_all", [(4, 5)], lambda s, a: a.max()),
    ("min_axis0", [(4, 5, 6)], lambda s, a: a.min(0)),
    ("min_axes_keep", [(4, 5, 6)], lambda s, a: a.min(0, 2, keepdims=True)),
    ("logsumexp_last", [(6, 10)], lambda s, a: s.logsumexp(a, -1)),
    ("logsumexp_keep", [(6, 10)], lambda s, a: s.logsumexp(a, 1, keepdims=True)),
    ("logsumexp_all", [(3, 4)], lambda s, a: s.logsumexp(a)),
    ("log_exp", [(3, 5)], lambda s, a: s.log(s.exp(a) + 1.0)),
    # ---- shape ops (tensor.pyx:1629, 1863-1977) ---------------------------------------------
    ("broadcast_to", [(1, 5)], lambda s, a: a.broadcast_to(4, 3, 5)),
    ("broadcast_to_tuple", [(3, 1)], lambda s, a: a.broadcast_to((3, 6))),
    ("reshape", [(4, 6)], lambda s, a: a.reshape(2, 12)),
    ("reshape_of_transpose", [(4, 6)], lambda s, a: a.T.reshape(3, 8)),
    ("permute", [(2, 3, 4)], lambda s, a: a.permute(2, 0, 1)),
    ("permute_then_mul", [(2, 3, 4), (4, 2, 3)], lambda s, a, b: a.permute(2, 0, 1) * b),
    ("transpose_reverses_all_axes", [(2, 3, 4)], lambda s, a: a.transpose()),
    ("T_matmul", [(5, 3), (5, 4)], lambda s, a, b: a.T @ b),
    # ---- indexing (tensor.pyx:1979, 922) -------------------------------------------------------
    ("getitem_int", [(4, 6)], lambda s, a: a[2]),
    ("getitem_neg_int", [(4, 6)], lambda s, a: a[-1]),
    ("getitem_slice", [(6, 8)], lambda s, a: a[1:5:2]),
    ("getitem_tuple", [(6, 8)], lambda s, a: a[1:5, ::3]),
    ("getitem_int_slice", [(4, 5, 6)], lambda s, a: a[1, :, 2:5]),
    ("getitem_chain", [(4, 5, 6)], lambda s, a: a[1:][0] * 2.0),
    # ---- matmul (tensor.pyx:1824) -------------------------------------------------------------
    ("matmul_2d", [(7, 5), (5, 3)], lambda s, a, b: a @ b),
    ("matmul_batched", [(2, 4, 5), (2, 5, 3)], lambda s, a, b: a @ b),
    ("matmul_bcast_batch", [(3, 2, 4, 5), (5, 6)], lambda s, a, b: a @ b),
    ("matmul_bcast_batch_lhs", [(4, 5), (3, 5, 2)], lambda s, a, b: a @ b),
    # ---- composites -----------------------------------------------------------------------------
    ("softmax", [(6, 10)], lambda s, a: s.exp(a - a.max(-1, keepdims=True)) /
        s.exp(a - a.max(-1, keepdims=True)).sum(-1, keepdims=True)),
    ("reuse_of_a_node", [(4, 4)], lambda s, a: (a * a + a) * a - a.sum(0)),
    ("diamond", [(3, 4), (4,)], lambda s, a, b: ((a + b) * (a - b)).sum(1) + (a * b).mean(1)),
    ("stack_no_grad", [(3, 4), (3, 4)], lambda s, a, b: s.stack([a, b], axis=1)),
    # ---- adjoint aliasing: add / sub backward hand the SAME adjoint to both inputs and reshape /
    # transpose / same-shape sum backward return VIEWS of it; an input with a second consumer must not
    # accumulate into storage another gradient still reads (autodiff.pyx:30-41, backward.pyx:60-86)
    ("alias_reshape_add_multi_consumer", [(2, 3), (6,)],
        lambda s, x, w: ((x.reshape(6) + w) * 3.0).sum() + (x * 2.0).sum()),
    ("alias_transpose_add_multi_consumer", [(3, 4), (4, 3)],
        lambda s, a, b: (a.T + b) * 2.0 + (a * a).T),
    ("alias_sub_permute_multi_consumer", [(2, 3, 4), (4, 2, 3)],
        lambda s, a, b: (b - a.permute(2, 0, 1)) * 1.5 + (a * 0.5).permute(2, 0, 1) + b * b),
    ("alias_same_shape_sum_multi_consumer", [(1, 4), (1, 4)],
        lambda s, a, b: (a.sum(0, keepdims=True) + b) * 3.0 + a * 2.0),
    ("alias_add_both_inputs_reused", [(3, 4), (3, 4)],
        lambda s, a, b: (a + b) * (a - b) + (a + b) + a * 2.0 + b * 3.0),
]

# forward-only cases: integer / bool results, never differentiable (tensor.pyx:812-920, 2050-2338)
INT_CASES = [
    ("argmax_last", [(6, 10)], lambda s, a: a.argmax(-1)),
    ("argmax_axis0_keep", [(6, 10)], lambda s, a: a.argmax(0, keepdims=True)),
    ("argmax_all", [(3, 4)], lambda s, a: a.argmax()),
    ("argmin_axis1", [(6, 10)], lambda s, a: a.argmin(1)),
    ("eq", [(4, 5), (4, 5)], lambda s, a, b: (a * 0 + 1.0) == (b * 0 + 1.0)),
    ("ne_scalar", [(4, 5)], lambda s, a: a != 0.0),
    ("gt", [(4, 5), (5,)], lambda s, a, b: a > b),
    ("ge_scalar", [(4, 5)], lambda s, a: a >= 0.25),
    ("lt", [(4, 5), (4, 1)], lambda s, a, b: a < b),
    ("le_is_ge_quirk_q6", [(4, 5), (4, 5)], lambda s, a, b: a <= b),
    ("one_hot", [(7,)], lambda s, a: s.one_hot(_t(s, np.array([1, 0, 3, 2, 2, 0, 3], "uint8")), 4)),
    ("one_hot_infer_classes", [(7,)], lambda s, a: s.one_hot(_t(s, np.array([1, 0, 3, 2, 2, 0, 5], "int32")))),
]


def make_inputs(shapes, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(shp).astype("float32") for shp in shapes]


# ---- dtype semantics: promotion table (soket/dtype.pyx), scalar typing (dtype.pyx:157-167), casts,
# creation functions (soket/tensor/creation.pyx).  Forward only; the result's dtype TAG and values are
# compared exactly (values within 1e-5 for float results).
def _ints(shape, dtype, seed):
    return np.random.default_rng(seed).integers(1, 9, shape).astype(dtype)


def _try(thunk):
    """The result, or the exception type when the expression is rejected (e.g. the reference's
    true-divide on integer result dtypes: numpy has no such loop -> TypeError)."""
    try:
        return thunk()
    except Exception as e:
        return "raises " + type(e).__name__


def _mixed(op):
    def run(s, a):
        out = []
        pairs = [("int32", "float32"), ("uint8", "int32"), ("int64", "float32"), ("uint8", "float64"),
                 ("int8", "uint8"), ("int64", "uint64"), ("float16", "float32"), ("int32", "int64"),
                 ("bool", "int32"), ("bool", "float32"), ("float32", "float64"), ("uint16", "int16")]
        for i, (da, db) in enumerate(pairs):
            x = _t(s, _ints((2, 3), da, i) if da != "bool" else (_ints((2, 3), "int8", i) > 4))
            y = _t(s, _ints((2, 3), db, 100 + i) if db != "bool" else (_ints((2, 3), "int8", 100 + i) > 4))
            out.append(_try(lambda: op(x, y)))
        return out
    return run


def _scalars(s, a):
    out = []
    for i, dt in enumerate(["int32", "uint8", "int64", "float32", "float64", "float16", "bool"]):
        x = _t(s, _ints((2, 3), dt, i) if dt != "bool" else (_ints((2, 3), "int8", i) > 4))
        out += [_try(f) for f in (lambda: x + 2, lambda: x * 2.5, lambda: x - True, lambda: 3 - x, lambda: x / 2,
                                  lambda: 2.0 / (x + 1), lambda: x ** 2, lambda: x * -1)]
    return out


def _casts(s, a):
    base = _t(s, (np.arange(12, dtype="float32").reshape(3, 4) - 4.5) * 1.5)
    out = []
    for name in ["float16", "float32", "float64", "int8", "uint8", "int16", "int32", "int64", "bool"]:
        out.append(s.Tensor(base, dtype=getattr(s, name)))
    ints = _t(s, np.arange(-3, 9, dtype="int32").reshape(3, 4))
    out += [s.Tensor(ints, dtype=s.float32), s.Tensor(ints, dtype=s.uint8), s.Tensor(ints, dtype=s.bool),
            s.Tensor([1, 2, 3]), s.Tensor([1.5, 2.5]), s.Tensor(7), s.Tensor(2.5), s.Tensor(True),
            s.Tensor([[1, 2], [3, 4]], dtype=s.float64), base.copy(), base.detach(), ints.sum(), ints.mean(),
            ints.max(), (ints > 2).sum(), (ints > 2).mean(dtype=s.float32), ints.sum(0, dtype=s.float32)]
    return out


def _creation(s, a):
    like = _t(s, np.zeros((2, 5), "int32"))
    return [s.zeros(2, 3), s.ones(4), s.zeros((2, 2), dtype=s.int32), s.ones(2, 2, dtype=s.float64),
            s.full(2, 3, fill=2.5), s.full((3,), fill=7, dtype=s.int64), s.zeros_like(like), s.one_like(like),
            s.zeros_like(like, dtype=s.float32), s.empty(2, 3) * 0.0,
            s.rand(3, 4) * 0.0, s.randn(2, 2, dtype=s.float64) * 0.0, s.randb(5, p=1.0), s.randb(2, 3, p=0.0)]


def _matmul_dtypes(s, a):
    """np.matmul's result dtype follows the promoted Tensor dtype (forward.pyx:172-178): float64 and
    integer operands keep their precision / type instead of being computed in float32."""
    rng = np.random.default_rng(7)
    f64a, f64b = rng.standard_normal((5, 7)), rng.standard_normal((7, 3))
    big = np.array([[2 ** 20 + 1, 3], [5, 2 ** 21 + 7]], "int64")
    out = [_t(s, f64a) @ _t(s, f64b),
           _t(s, f64a.astype("float32")) @ _t(s, f64b),
           _t(s, _ints((4, 6), "int32", 1)) @ _t(s, _ints((6, 2), "int32", 2)),
           _t(s, big) @ _t(s, big),
           _t(s, _ints((4, 6), "int32", 3)) @ _t(s, f64b[:6].astype("float32")),
           _t(s, _ints((2, 3, 4), "int16", 4).astype("float64")) @ _t(s, f64a[:4]),
           _t(s, f64a).T @ _t(s, f64a)]
    return [_try(lambda o=o: o) for o in out]


MULTI_CASES = [
    ("matmul_dtypes", _matmul_dtypes),
    ("mixed_add", _mixed(lambda x, y: x + y)),
    ("mixed_sub", _mixed(lambda x, y: x - y)),
    ("mixed_mul", _mixed(lambda x, y: x * y)),
    ("mixed_div", _mixed(lambda x, y: x / y)),
    ("mixed_eq", _mixed(lambda x, y: x == y)),
    ("mixed_gt", _mixed(lambda x, y: x > y)),
    ("scalar_operands", _scalars),
    ("casts_and_constructors", _casts),
    ("creation_functions", _creation),
]
