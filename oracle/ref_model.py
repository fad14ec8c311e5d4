"""oracle/ref_model.py -- drive the BUILT reference (oracle/_ref) as an oracle.

TEST INFRASTRUCTURE ONLY.  Builds the north-star model of
examples/mlp_resnet/model.py:17-58 out of the reference's own ``soket.nn``
modules (or any API-compatible ``nn`` namespace, e.g. ``soket_b200.engine.nn``),
in the `self.fn`-retaining variant that makes the inner layers visible to
``parameters()`` / ``modules()``, and maps parameters to the
flat names used by ``oracle/soket_np.py``.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def import_reference(install_compat=True):
    """Import the built reference package; returns the `soket` module or None.
    install_compat=False keeps soket_b200 out of the process (bench.py --impl reference: the
    reference arm must not load the product).

    When soket_b200 is importable it is first registered under the module name the
    reference imports for its GPU backend (soket_b200.compat.install), so the same
    process can run the reference on its CPU device (NumPy -- the oracle) AND on
    `soket.gpu()` (the sm_100a kernels -- the drop-in test)."""
    if not os.path.exists(os.path.join(REF_DIR, "soket", "__init__.py")):
        return None
    if install_compat and "soket" not in sys.modules:
        try:
            import soket_b200.compat as compat
            compat.install()
        except Exception:
            pass
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import soket  # noqa: F401
    return soket


def build_model(nn, dim, hidden, num_blocks, num_classes, norm="layer", drop_prob=0.0, retain_fn=True):
    """Same composition as model.py:17-58.  `retain_fn=False` is the verbatim model
    (inner layers invisible to parameters(): quirk Q1)."""
    norm_cls = nn.LayerNorm if norm == "layer" else nn.BatchNorm1d

    class ResidualBlock(nn.Sequential):
        def __init__(self, d, h):
            fn = nn.Sequential(
                nn.Linear(d, h), norm_cls(h), nn.ReLU(), nn.Dropout(p=drop_prob),
                nn.Linear(h, h), norm_cls(h))
            super().__init__(nn.Residual(fn), nn.ReLU())
            if retain_fn:
                self.fn = fn

    class MLPResNet(nn.Sequential):
        def __init__(self):
            blocks = [ResidualBlock(hidden, hidden) for _ in range(num_blocks)]
            super().__init__(nn.Linear(dim, hidden), nn.ReLU(), *blocks, nn.Linear(hidden, num_classes))

    return MLPResNet()


def named_parameters(model, num_blocks):
    """Flat name -> Tensor, names as in oracle.soket_np.MLPResNet (needs retain_fn)."""
    mods = list(model.modules())
    lin0 = None
    out = {}
    # model._storage order: Linear, ReLU, blocks..., Linear -- reach them via modules()
    linears = [m for m in mods if type(m).__name__ == "Linear"]
    norms = [m for m in mods if type(m).__name__ in ("LayerNorm", "BatchNorm1d")]
    # modules() walks the model's storage in order; each block exposes its inner
    # layers through its `fn` attribute: lin0, blk0.lin1, blk0.lin2, ..., out
    assert len(linears) == 2 + 2 * num_blocks, len(linears)
    lin0, last = linears[0], linears[-1]
    inner = linears[1:-1]
    out["lin0.W"], out["lin0.b"] = lin0.weight, lin0.bias
    for i in range(num_blocks):
        l1, l2 = inner[2 * i], inner[2 * i + 1]
        n1, n2 = norms[2 * i], norms[2 * i + 1]
        out[f"blk{i}.lin1.W"], out[f"blk{i}.lin1.b"] = l1.weight, l1.bias
        g1, b1 = list(n1.parameters())
        out[f"blk{i}.n1.g"], out[f"blk{i}.n1.b"] = g1, b1
        out[f"blk{i}.lin2.W"], out[f"blk{i}.lin2.b"] = l2.weight, l2.bias
        g2, b2 = list(n2.parameters())
        out[f"blk{i}.n2.g"], out[f"blk{i}.n2.b"] = g2, b2
    out["out.W"], out["out.b"] = last.weight, last.bias
    return out


def to_numpy(soket, t):
    """A reference Tensor (CPU device) as a NumPy array: the reference has no .numpy(); write it
    through Tensor.__setitem__ into a from_numpy view (tensor.pyx:484,1092)."""
    import numpy as np
    if len(t.shape) == 0:
        return np.array(t.item(), dtype=str(t.dtype))
    buf = np.zeros(t.shape, dtype=str(t.dtype))
    view = soket.Tensor.from_numpy(buf)
    view[tuple(slice(None) for _ in t.shape)] = t
    return buf


# --- ReLU sign patterns: comparing both backends at one and the same point of the derivative ---------

def device_relu_signs(model, X, blocks):
    """The sign pattern of every ReLU input of the FUSED forward (y > 0 <=> pre-activation > 0), read
    off the fused kernels' own outputs through Sequential._run prefixes."""
    import soket_b200.api as soket                # the device side of the comparison (tests / DP parity worker)
    signs = []
    h = model._run(soket.Tensor(X), 2)                       # relu(lin0(X)): GEMM with bias+ReLU epilogue
    signs.append(h.numpy() > 0)
    mods = list(model._storage)
    for i in range(blocks):
        blk = mods[2 + i]
        r1 = blk.fn._run(h, 3)                               # relu(LN1(lin1(h))): LN kernel, ReLU epilogue
        signs.append(r1.numpy() > 0)
        h = blk(h)                                           # relu(h + LN2(...)): LN kernel, residual epilogue
        signs.append(h.numpy() > 0)
    return signs


def oracle_relu_signs(om, blocks):
    T = om.tape
    out = [T["lin0.pre"] > 0]
    for i in range(blocks):
        out += [T[f"blk{i}"]["relu1.in"] > 0, T[f"blk{i}"]["relu2.in"] > 0]
    return out


def reconcile_relu_masks(om, dev_signs, blocks):
    """The device's ReLU sign patterns as the mask list O.MLPResNet.backward takes, after checking
    that they differ from the oracle's only where the oracle's pre-activation is within rounding
    distance of zero.  Returns (masks, number of differing entries)."""
    T = om.tape
    pre = [T["lin0.pre"]]
    for i in range(blocks):
        pre += [T[f"blk{i}"]["relu1.in"], T[f"blk{i}"]["relu2.in"]]
    flips = 0
    for z, dev in zip(pre, dev_signs):
        diff = dev != (z > 0)
        n = int(diff.sum())
        if n:
            flips += n
            # both sides computed the same sum of O(1) terms to ~1e-7 relative of the terms
            assert float(np.abs(z[diff]).max()) <= 2e-6 * max(1.0, float(np.abs(z).max())), float(np.abs(z[diff]).max())
    return [np.asarray(d, "float32") for d in dev_signs], flips
