"""Pre-split GEMM operands (sk_split_f16 / sk_gemm_f16x3 / sk_layernorm_*_ex; include/soket_b200.h):
the resident path rewrites each matrix of a training step as fp16 hi / lo ONCE, with one power-of-two
scale per matrix, instead of inside every GEMM call.

Checked here: the GEMM on pre-split operands against float64 with the fp32-matmul bound
|err| <= 1e-5 (|a| @ |b|) for every operand orientation (the matmuls of forward.pyx:172-178 and
backward.pyx:720-736: x @ w, adj @ w.T, x.T @ adj), bit-identical to the per-call split of round 1,
the accumulate epilogue (autodiff.pyx:30-41 in place), the splits / |max| the LayerNorm kernels emit,
cache invalidation on every kind of in-place write, and that a model trains to the same numbers with
the path on and off.
"""
import numpy as np
import pytest

from oracle import ref_model, soket_np as O

pytestmark = pytest.mark.gpu


def f64_bound(a, b):
    return a.astype(np.float64) @ b.astype(np.float64), 1e-5 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))


@pytest.mark.parametrize("M,K,N", [(512, 384, 256), (784, 1024, 512), (1024, 512, 4096), (300, 264, 136)])
def test_gemm_on_presplit_operands_all_orientations(sk, M, K, N):
    rng = np.random.default_rng(M + K + N)
    a = rng.standard_normal((M, K)).astype("float32")
    b = (rng.standard_normal((K, N)) * 0.05).astype("float32")
    want, bound = f64_bound(a, b)
    sa, sb = sk.split_f16(sk.array(a)), sk.split_f16(sk.array(b))
    sat, sbt = sk.split_f16(sk.array(np.ascontiguousarray(a.T))), sk.split_f16(sk.array(np.ascontiguousarray(b.T)))
    base = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b)))               # round-1 path: per-row / per-column scales
    for ta, tb, x, y in [(False, False, sa, sb), (True, False, sat, sb), (False, True, sa, sbt), (True, True, sat, sbt)]:
        got = sk.asnumpy(sk.gemm_split(x, ta, y, tb))
        assert got.shape == (M, N)
        assert np.all(np.abs(got - want) <= bound), (ta, tb, float((np.abs(got - want) / bound).max()))
        # a power-of-two scale only moves exponents: same products, same accumulation order -> same bits
        assert np.array_equal(got, base), (ta, tb, float(np.abs(got - base).max()))


def test_split_reconstructs_and_pads(sk):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((300, 204)) * np.exp(rng.standard_normal((300, 1)) * 3)).astype("float32")
    m = sk.split_f16(sk.array(x))
    assert (m.rows, m.cols, m.ld) == (300, 204, 208)
    hi, lo, sc = sk.asnumpy(m.hi).astype(np.float64), sk.asnumpy(m.lo).astype(np.float64), sk.asnumpy(m.scale)
    amax = float(np.abs(x).max())
    assert sc[2] == np.float32(amax) and sc[0] * sc[1] == 1.0
    assert 2.0 ** 14 <= amax * sc[0] < 2.0 ** 15                           # top of the fp16 range, one bit of headroom
    assert np.all(hi[:, 204:] == 0) and np.all(lo[:, 204:] == 0)          # K padding contributes nothing
    rec = (hi[:, :204] + lo[:, :204]) / sc[0]
    # 22 significant bits, or an absolute 2^-25 of a scaled unit for elements far below the maximum
    assert np.all(np.abs(rec - x) <= 2.0 ** -22 * np.abs(x) + 2.0 ** -25 / sc[0])
    with pytest.raises(ValueError):
        sk.split_f16(sk.array(x[:, :203]))                                  # rows must start 16-byte aligned
    # the bias gradient rides along
    m2, cs = sk.split_f16(sk.array(x), True)
    assert np.array_equal(sk.asnumpy(m2.hi), sk.asnumpy(m.hi))
    assert np.allclose(sk.asnumpy(cs), x.astype(np.float64).sum(0), rtol=1e-5, atol=1e-5 * np.abs(x).sum(0).max())


def test_accumulate_epilogue_equals_separate_add(sk):
    from soket_b200 import _fused as F
    rng = np.random.default_rng(1)
    a = rng.standard_normal((512, 256)).astype("float32")
    b = rng.standard_normal((256, 384)).astype("float32")
    c0 = rng.standard_normal((512, 384)).astype("float32")
    sa, sb = sk.split_f16(sk.array(a)), sk.split_f16(sk.array(b))
    prod = sk.gemm_split(sa, False, sb, False)
    want = F.accumulate_(sk.array(c0), prod)                               # autodiff.pyx:30-41: c0 + prod
    out = sk.array(c0)
    got = sk.gemm_split(sa, False, sb, False, None, False, out, True)
    assert got is out
    assert np.array_equal(sk.asnumpy(got), sk.asnumpy(want))


@pytest.mark.parametrize("rows,cols", [(512, 512), (1024, 4096), (300, 1024)])
@pytest.mark.parametrize("variant", ["plain", "relu", "residual", "dropout"])
def test_layernorm_emits_the_split_of_its_output(sk, rows, cols, variant):
    """The split a LayerNorm kernel writes next to y reconstructs y, its scale comes from the
    a-priori bound max|gamma| sqrt(C) + max|beta| (+ the residual's bound, x 1/keep), and the GEMM fed
    with it meets the fp32-matmul bound."""
    from soket_b200 import _fused as F
    rng = np.random.default_rng(rows + cols)
    x = (rng.standard_normal((rows, cols)) * 2 + 0.5).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    be = (rng.standard_normal(cols) * 0.1).astype("float32")
    w = (rng.standard_normal((cols, 256)) * np.sqrt(2.0 / cols)).astype("float32")
    xd, gd, bd = sk.array(x), sk.array(g), sk.array(be)
    bound = float(np.abs(g).max()) * np.sqrt(cols) + float(np.abs(be).max())
    if variant == "residual":
        res = rng.standard_normal((rows, cols)).astype("float32")
        rd = sk.array(res)
        rs = sk.split_f16(rd)
        y, _, _ = F.layernorm_fwd(xd, gd, bd, rd, 1e-5, True, True, rs)
        ref, _, _ = F.layernorm_fwd(xd, gd, bd, rd, 1e-5, True)
        bound += float(np.abs(res).max())
    elif variant == "dropout":
        sk.random.seed(5)
        y, _, _, seed = F.layernorm_dropout_fwd(xd, gd, bd, 1e-5, True, 0.9, True)
        ref = None
        bound /= 0.9
    else:
        y, _, _ = F.layernorm_fwd(xd, gd, bd, None, 1e-5, variant == "relu", True)
        ref, _, _ = F.layernorm_fwd(xd, gd, bd, None, 1e-5, variant == "relu")
    m = y._meta
    assert type(m).__name__ == "SplitMat" and (m.rows, m.cols) == (rows, cols)
    yn = sk.asnumpy(y)
    if ref is not None:
        assert np.array_equal(yn, sk.asnumpy(ref))                         # the fp32 result is untouched
    sc = sk.asnumpy(m.scale)
    assert abs(sc[2] - bound * 1.0001) <= 1e-5 * bound and float(np.abs(yn).max()) <= sc[2]
    assert 2.0 ** 14 <= sc[2] * sc[0] < 2.0 ** 15
    rec = (sk.asnumpy(m.hi).astype(np.float64) + sk.asnumpy(m.lo).astype(np.float64)) / sc[0]
    assert np.all(np.abs(rec - yn) <= 2.0 ** -22 * np.abs(yn) + 2.0 ** -25 / sc[0])
    assert sk.get_split(y) is m                                            # consumed as is by the next Linear
    got = sk.asnumpy(sk.gemm_split(m, False, sk.split_f16(sk.array(w)), False))
    want, bnd = f64_bound(yn, w)
    assert np.all(np.abs(got - want) <= bnd), float((np.abs(got - want) / bnd).max())


@pytest.mark.parametrize("rows,cols", [(512, 512), (1024, 4096)])
def test_layernorm_backward_emits_absmax_of_dx(sk, rows, cols):
    from soket_b200 import _fused as F
    rng = np.random.default_rng(3)
    x = rng.standard_normal((rows, cols)).astype("float32")
    g = (rng.random(cols) + 0.5).astype("float32")
    be = (rng.standard_normal(cols) * 0.1).astype("float32")
    adj = (rng.standard_normal((rows, cols)) * np.exp(rng.standard_normal((rows, 1)) * 2)).astype("float32")
    xd, gd, bd = sk.array(x), sk.array(g), sk.array(be)
    _, mean, rstd = F.layernorm_fwd(xd, gd, bd, None, 1e-5, True)
    dx = F.layernorm_bwd(sk.array(adj), xd, gd, bd, mean, rstd, None, 1, False, True, None, None, True)[0]
    am = dx._meta
    assert type(am).__name__ == "AbsMax"
    dxn = sk.asnumpy(dx)
    assert sk.asnumpy(am.word).view(np.float32)[0] == np.abs(dxn).max()
    m = sk.split_f16(dx)                                                    # takes its scale from the word
    assert sk.asnumpy(m.scale)[2] == np.abs(dxn).max()
    assert np.array_equal(sk.asnumpy(m.hi), sk.asnumpy(sk.split_f16(sk.array(dxn)).hi))


def test_split_cache_follows_in_place_writes(sk):
    """A cached split is dropped by anything that rewrites the storage: __setitem__ through any view,
    fill, accumulate_, the optimizer kernels, a graph replay."""
    from soket_b200 import _fused as F
    rng = np.random.default_rng(2)
    w = sk.array(rng.standard_normal((256, 256)).astype("float32"))
    m = sk.get_split(w)
    assert sk.get_split(w) is m
    w[0, 0] = 3.0
    m2 = sk.get_split(w)
    assert m2 is not m and sk.get_split(w) is m2
    w.reshape(-1)[5:9] = sk.array(np.ones(4, "float32"))                    # through another view of the buffer
    m3 = sk.get_split(w)
    assert m3 is not m2
    F.accumulate_(w, sk.array(np.ones((256, 256), "float32")))
    m4 = sk.get_split(w)
    assert m4 is not m3
    F.sgd_step([w], [sk.array(np.ones((256, 256), "float32"))], 0.1)
    m5 = sk.get_split(w)
    assert m5 is not m4
    want = sk.asnumpy(sk.split_f16(sk.array(sk.asnumpy(w))).hi)
    assert np.array_equal(sk.asnumpy(m5.hi), want)                          # and the fresh one is of the NEW contents
    w.fill(0.5)
    assert sk.get_split(w) is not m5


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_model_trains_to_the_same_numbers_with_and_without_presplit(sk, opt):
    """MLPResNet(784, 512, 2 blocks) at batch 512, five steps: the pre-split path (weights split once
    per step, activations split by the LayerNorm kernels, adjoints once, dX accumulating into the
    residual branch's adjoint) against the per-call splits of round 1."""
    import soket_b200.api as soket
    from soket_b200 import engine as E, nn
    from soket_b200.optim import SGD, Adam
    dim, hidden, nb, C, Bn = 784, 512, 2, 10, 512
    runs = {}
    for mode in (True, False):
        E.set_presplit(mode)
        try:
            rng = np.random.default_rng(0)
            model = ref_model.build_model(nn, dim, hidden, nb, C, norm="layer", drop_prob=0.0)
            named = ref_model.named_parameters(model, nb)
            for k, t in named.items():
                shp = tuple(t.shape)
                v = rng.standard_normal(shp) * (np.sqrt(2.0 / shp[0]) if k.endswith(".W") else 0.1)
                t.data = soket.Tensor((v + (1.0 if k.endswith(".g") else 0.0)).astype("float32"))
            o = SGD(model.parameters(), lr=0.01) if opt == "sgd" else Adam(model.parameters(), lr=0.001)
            crit = nn.SoftmaxCrossEntropyLoss()
            losses = []
            sk.profile_reset(); sk.profile_enable(True)
            for s in range(5):
                X = rng.random((Bn, dim), dtype=np.float32)
                y = rng.integers(0, C, Bn).astype(np.uint8)
                loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
                loss.backward()
                if s == 0:
                    g0 = {k: t.grad.numpy() for k, t in named.items()}
                o.step()
                losses.append(loss.item())
            sk.profile_enable(False)
            runs[mode] = (losses, g0, {k: t.numpy() for k, t in named.items()}, sk.profile_collect())
        finally:
            E.set_presplit(True)
    (l1, g1, p1, prof1), (l0, g0, p0, prof0) = runs[True], runs[False]
    # fewer and cheaper operand passes, same GEMM count
    assert prof1["gemm_tc"]["launches"] == prof0["gemm_tc"]["launches"]
    assert prof1["gemm_prep"]["work"] < 0.7 * prof0["gemm_prep"]["work"], (prof1["gemm_prep"], prof0["gemm_prep"])
    assert "ewise" not in prof1 or prof1["ewise"]["launches"] < prof0["ewise"]["launches"]   # no accumulate pass
    assert np.allclose(l1, l0, rtol=1e-6, atol=0), (l1, l0)
    for k in g1:
        scale = max(float(np.abs(g0[k]).max()), 1e-30)
        assert float(np.abs(g1[k] - g0[k]).max()) <= 2e-6 * scale, k
    if opt == "sgd":
        for k in p1:
            assert float(np.abs(p1[k] - p0[k]).max()) <= 1e-6 * max(float(np.abs(p0[k]).max()), 1e-30), k


def test_adam_rewrites_the_weight_split_in_its_own_kernel(sk):
    """sk_adam_step_split: from the second step on, the optimizer kernel leaves the new weights' hi / lo
    next to the weights themselves (scale chosen a priori from max |w_old| + lr * ratio bound), the
    split stays bound to the weight array, reconstructs it to 22 bits, and feeds the GEMM the very
    bits a fresh sk_split_f16 pass would; the update itself is unchanged (bit-identical weights with
    the fusion off)."""
    from soket_b200 import _fused as F
    from soket_b200 import optim as OPT
    rng = np.random.default_rng(5)
    w0 = (rng.standard_normal((512, 384)) * 0.05).astype("float32")
    x = sk.array(rng.standard_normal((256, 512)).astype("float32"))
    sx = sk.split_f16(x)
    grads = [(rng.standard_normal((512, 384)) * 10.0 ** rng.integers(-6, 1)).astype("float32") for _ in range(6)]

    def run(fused):
        w = sk.array(w0)
        m, v = sk.empty((512, 384), "float32"), sk.empty((512, 384), "float32")
        b1, b2, lr = 0.9, 0.999, 1e-3
        b1t, b2t = b1, b2
        outs, launches = [], []
        for t, g in enumerate(grads, 1):
            ub = abs(lr) * OPT.adam_ratio_bound(b1, b2, t) * 1.0001 if fused else -1.0
            F.adam_step([w], [sk.array(g)], [m], [v], lr, b1, b2, 1e-8, 0.0, 1.0 - b1t, 1.0 - b2t, t == 1, 1.0, None, True, ub)
            b1t *= b1; b2t *= b2
            n0 = sk.launch_count()
            sw = sk.get_split(w)
            launches.append(sk.launch_count() - n0)
            outs.append((sk.asnumpy(w), sk.asnumpy(sk.gemm_split(sx, False, sw, False)), sk.asnumpy(sw.hi).astype(np.float64),
                         sk.asnumpy(sw.lo).astype(np.float64), sk.asnumpy(sw.scale)))
        return outs, launches
    fused, lf = run(True)
    plain, lp = run(False)
    assert lf[0] >= 1 and all(n == 0 for n in lf[1:]), lf     # step 1 has no split to refresh yet; then no pass at all
    assert all(n >= 1 for n in lp), lp
    for (wf, yf, hi, lo, sc), (wp, yp, _, _, scp) in zip(fused, plain):
        assert np.array_equal(wf, wp)                                       # the update is the same arithmetic
        assert np.array_equal(yf, yp)                                       # a power-of-two scale only moves exponents
        amax = float(np.abs(wf).max())
        assert sc[0] * sc[1] == 1.0 and amax <= sc[2] and amax * sc[0] < 2.0 ** 15   # the a-priori bound held
        assert sc[2] <= 1.5 * amax + 1e-2                                   # and is not wildly loose (lr * 7.3 of slack)
        rec = (hi + lo) / sc[0]
        assert np.all(np.abs(rec - wf) <= 2.0 ** -22 * np.abs(wf) + 2.0 ** -25 / sc[0])
