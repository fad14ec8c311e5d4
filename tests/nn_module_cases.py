"""nn-module cases shared by the reference (oracle/_ref, CPU) and soket_b200 (GPU): each case
is fn(soket, nn, rng) -> dict of named results (Tensors, strings or numbers).  Parameters are
overwritten with seeded values through `p.data = Tensor(...)` (what soket/nn/init.py does), so
both backends start from identical weights; inputs with requires_grad report their gradients.
SURVEY.md section 8(a) composite ops + 8(f)-3 (BatchNorm2d/3d, channels cross-entropy, ...)."""
from collections import OrderedDict

import numpy as np


def seed_params(soket, module, rng, scale=0.5):
    for p in module.parameters():
        shape = tuple(p.shape)
        p.data = soket.Tensor((scale * rng.standard_normal(shape)).astype("float32") + (1.0 if len(shape) == 1 else 0.0))


def run(soket, module, rng, x_shape, extra=None, weight_seed=99, y=None):
    """forward, sum(out * w).backward(); returns out, input grad, parameter grads, str(module)."""
    x = soket.Tensor(rng.standard_normal(x_shape).astype("float32"), requires_grad=True)
    out = module(x) if y is None else module(x, y)
    res = {"out": out}
    try:
        res["str"] = str(module)
    except Exception as e:          # the reference's BatchNorm __str__ reads an attribute it never sets
        res["str"] = "raises " + type(e).__name__
    if out.requires_grad:
        w = soket.Tensor(np.random.default_rng(weight_seed).standard_normal(tuple(out.shape)).astype("float32"))
        (out * w).sum().backward()
        res["dx"] = x.grad
        for i, p in enumerate(module.parameters()):
            if p.grad is not None:
                res[f"dp{i}"] = p.grad
    if extra:
        res.update(extra(module))
    return res


def linear_2d(soket, nn, rng):
    m = nn.Linear(6, 4)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (5, 6))


def linear_no_bias_3d(soket, nn, rng):
    m = nn.Linear(6, 4, bias=False)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (2, 5, 6))


def linear_relu_sequential(soket, nn, rng):
    m = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 3))
    seed_params(soket, m, rng)
    return run(soket, m, rng, (7, 6))


def sequential_ordered_dict_append_identity(soket, nn, rng):
    m = nn.Sequential(OrderedDict([("fc", nn.Linear(4, 4)), ("act", nn.ReLU())]))
    m.append(nn.Linear(4, 2))
    seed_params(soket, m, rng)
    res = run(soket, m, rng, (3, 4))
    res["n_params"] = len(list(m.parameters()))
    res["n_modules"] = len(list(m.modules()))
    return res


def identity_module(soket, nn, rng):
    return run(soket, nn.Identity(), rng, (3, 4))


def residual_block(soket, nn, rng):
    fn = nn.Sequential(nn.Linear(8, 8), nn.LayerNorm(8), nn.ReLU(), nn.Linear(8, 8), nn.LayerNorm(8))
    m = nn.Sequential(nn.Residual(fn), nn.ReLU())
    seed_params(soket, fn, rng)
    res = run(soket, m, rng, (6, 8))
    res["n_params_visible"] = len(list(m.parameters()))      # quirk Q1: 0
    return res


def layernorm_2d(soket, nn, rng):
    m = nn.LayerNorm(12)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (5, 12))


def layernorm_no_affine(soket, nn, rng):
    m = nn.LayerNorm(12, elementwise_affine=False)
    return run(soket, m, rng, (5, 12))


def layernorm_no_bias(soket, nn, rng):
    m = nn.LayerNorm(12, bias=False)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (5, 12))


def layernorm_3d_input(soket, nn, rng):
    m = nn.LayerNorm(8)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (2, 3, 8))


def layernorm_wide_fused_shape(soket, nn, rng):
    m = nn.LayerNorm(1024)
    seed_params(soket, m, rng, 0.1)
    return run(soket, m, rng, (9, 1024))


def _running(m):
    return {"running_mean": m.running_mean, "running_var": m.running_var}


def batchnorm1d_train_twice(soket, nn, rng):
    m = nn.BatchNorm1d(8)
    seed_params(soket, m, rng)
    m(soket.Tensor(rng.standard_normal((10, 8)).astype("float32")))
    return run(soket, m, rng, (10, 8))


def batchnorm1d_eval_flag_is_ignored_q4(soket, nn, rng):
    m = nn.BatchNorm1d(8)
    seed_params(soket, m, rng)
    m.train(False)
    return run(soket, m, rng, (10, 8))


def batchnorm1d_no_affine_no_stats(soket, nn, rng):
    m = nn.BatchNorm1d(8, affine=False, track_running_stats=False)
    return run(soket, m, rng, (10, 8))


def batchnorm2d(soket, nn, rng):
    m = nn.BatchNorm2d(3)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (2, 3, 4, 5))


def batchnorm3d(soket, nn, rng):
    m = nn.BatchNorm3d(2, momentum=0.3, eps=1e-3)
    seed_params(soket, m, rng)
    return run(soket, m, rng, (2, 2, 3, 2, 2))


def batchnorm_wide_fused_shape(soket, nn, rng):
    m = nn.BatchNorm1d(256)
    seed_params(soket, m, rng, 0.1)
    return run(soket, m, rng, (64, 256))


def dropout_eval_and_p0(soket, nn, rng):
    m = nn.Sequential(nn.Dropout(p=0.0), nn.Linear(4, 4))
    seed_params(soket, m, rng)
    res = run(soket, m, rng, (3, 4))
    d = nn.Dropout(p=0.7)
    d.train(False)
    x = soket.Tensor(rng.standard_normal((4, 4)).astype("float32"))
    res["eval_identity"] = d(x)
    res["dropout_str"] = str(d)
    return res


def _labels(soket, rng, shape, classes, dtype="uint8"):
    return soket.Tensor(rng.integers(0, classes, shape).astype(dtype))


def cross_entropy_mean(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(), rng, (6, 5), y=_labels(soket, rng, (6,), 5))


def cross_entropy_sum(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(reduction="sum"), rng, (6, 5), y=_labels(soket, rng, (6,), 5, "int32"))


def cross_entropy_none(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(reduction="none"), rng, (6, 5), y=_labels(soket, rng, (6,), 5, "int64"))


def cross_entropy_channels(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(), rng, (2, 5, 3), y=_labels(soket, rng, (2, 3), 5))


def cross_entropy_channels_2d(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(reduction="sum"), rng, (2, 4, 3, 2), y=_labels(soket, rng, (2, 3, 2), 4))


def cross_entropy_single_sample(soket, nn, rng):
    return run(soket, nn.SoftmaxCrossEntropyLoss(), rng, (5,), y=soket.Tensor(np.array(3, "uint8")))


def cross_entropy_large_logits(soket, nn, rng):
    m = nn.SoftmaxCrossEntropyLoss()
    x = soket.Tensor((60 * rng.standard_normal((8, 10))).astype("float32"), requires_grad=True)
    out = m(x, _labels(soket, rng, (8,), 10))
    out.backward()
    return {"out": out, "dx": x.grad}


def functional_forms(soket, nn, rng):
    import importlib
    fn = importlib.import_module(nn.__name__ + ".functional")
    x = soket.Tensor(rng.standard_normal((4, 6)).astype("float32"), requires_grad=True)
    g = soket.Tensor((1 + 0.1 * rng.standard_normal((6,))).astype("float32"), requires_grad=True)
    b = soket.Tensor((0.1 * rng.standard_normal((6,))).astype("float32"), requires_grad=True)
    out = fn.layer_norm(x, g, b, 1e-3) + fn.batch_norm(x, None, None, g, b, True, 0.1, 1e-3)
    w = soket.Tensor(rng.standard_normal((4, 6)).astype("float32"))
    (out * w).sum().backward()
    return {"out": out, "dx": x.grad, "dg": g.grad, "db": b.grad}


def module_tree_and_train_flag(soket, nn, rng):
    inner = nn.Sequential(nn.Linear(3, 3), nn.Dropout(p=0.5))
    m = nn.Sequential(nn.Linear(3, 3), nn.Residual(inner), nn.BatchNorm1d(3), nn.LayerNorm(3))
    m.train(False)
    return {"n_params": len(list(m.parameters())), "n_modules": len(list(m.modules())),
            "str": str(nn.Sequential(nn.Linear(3, 3), nn.Residual(inner), nn.LayerNorm(3), nn.ReLU(), nn.Dropout(p=0.25),
                                     nn.SoftmaxCrossEntropyLoss(reduction="sum"))),
            "shapes": str([tuple(p.shape) for p in m.parameters()])}


def error_messages(soket, nn, rng):
    out = {}
    x = soket.Tensor(rng.standard_normal((4, 5)).astype("float32"))
    for name, call in (
            ("ce_bad_reduction", lambda: nn.SoftmaxCrossEntropyLoss(reduction="avg")),
            ("ce_bad_target_shape", lambda: nn.SoftmaxCrossEntropyLoss()(x, soket.Tensor(np.zeros((3,), "uint8")))),
            ("seq_non_module", lambda: nn.Sequential(nn.ReLU(), 3)),
            ("matmul_inner_mismatch", lambda: x @ x),
            ("reshape_bad", lambda: x.reshape(3, 7)),
            ("broadcast_bad", lambda: x.broadcast_to(4, 6)),
            ("getitem_out_of_bounds", lambda: x[7]),
    ):
        try:
            call()
            out[name] = "no error"
        except Exception as e:      # the TYPE is the contract (SURVEY.md 8b errors); messages are compared loosely
            out[name] = type(e).__name__
    return out


CASES = [linear_2d, linear_no_bias_3d, linear_relu_sequential, sequential_ordered_dict_append_identity,
         identity_module, residual_block, layernorm_2d, layernorm_no_affine, layernorm_no_bias, layernorm_3d_input,
         layernorm_wide_fused_shape, batchnorm1d_train_twice, batchnorm1d_eval_flag_is_ignored_q4,
         batchnorm1d_no_affine_no_stats, batchnorm2d, batchnorm3d, batchnorm_wide_fused_shape,
         dropout_eval_and_p0, cross_entropy_mean, cross_entropy_sum, cross_entropy_none,
         cross_entropy_channels, cross_entropy_channels_2d, cross_entropy_single_sample,
         cross_entropy_large_logits, functional_forms, module_tree_and_train_flag, error_messages]
