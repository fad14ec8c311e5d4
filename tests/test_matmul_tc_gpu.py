"""GPU parity of the tcgen05 / TMEM / TMA GEMM against NumPy (np.matmul in float64 as the
exact reference, np.matmul fp32 as the oracle value) for every operand layout the
reference feeds it (forward.pyx:172-178: x @ w; backward.pyx:720-736: adj @ w.T and
x.T @ adj on .T VIEWS), ragged / tail shapes included.

Error norm of a dot product: |got - exact| <= tol * (|a| @ |b|).
  3xTF32 (the fp32 parity path): tol 1e-5 (north_star), typically ~1e-7
  single-pass TF32: tol 2e-3 (10-bit mantissa; offered for speed, NOT a parity path)
  bf16 inputs: exact products, fp32 accumulation -> tol 1e-5 against fp32 matmul of the
  bf16-rounded inputs
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 128), (256, 384, 512), (100, 784, 100), (1000, 100, 260), (129, 33, 257),
          (64, 64, 64), (2048, 1024, 512), (8, 8, 8), (300, 4, 300), (1, 128, 128), (130, 256, 1)]


def err_ratio(got, a, b):
    exact = a.astype(np.float64) @ b.astype(np.float64)
    bound = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    return (np.abs(got.astype(np.float64) - exact) / np.maximum(bound, 1e-30)).max()


def operands(sk, M, K, N, a_t, b_t, seed):
    """Device operands whose logical shapes are (M,K) and (K,N); *_t builds them as .T views."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, (M, K)).astype("float32")
    b = rng.uniform(-1, 1, (K, N)).astype("float32")
    da = sk.array(np.ascontiguousarray(a.T)).T if a_t else sk.array(a)
    db = sk.array(np.ascontiguousarray(b.T)).T if b_t else sk.array(b)
    assert da.shape == (M, K) and db.shape == (K, N)
    return a, b, da, db


@pytest.mark.parametrize("M,K,N", SHAPES)
@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
def test_tf32x3_all_layouts(sk, M, K, N, a_t, b_t):
    a, b, da, db = operands(sk, M, K, N, a_t, b_t, M + K + N)
    try:
        got = sk.asnumpy(sk.matmul(da, db, algo=sk.MM_TF32X3))
    except RuntimeError as e:
        # pitches that are not multiples of 16 bytes cannot be described to TMA; the
        # dispatcher (MM_AUTO) sends those to the FFMA kernel
        assert "tcgen05 path does not support" in str(e)
        lda = da.strides[0] if not a_t else da.strides[1]
        ldb = db.strides[0] if not b_t else db.strides[1]
        assert lda % 16 != 0 or ldb % 16 != 0
        got = sk.asnumpy(sk.matmul(da, db))
    assert got.shape == (M, N) and got.dtype == np.float32
    assert err_ratio(got, a, b) <= 1e-5
    assert err_ratio(got, a, b) <= 2e-6   # typical: a few 1e-7, like an FFMA kernel


@pytest.mark.parametrize("M,K,N", [(256, 384, 512), (132, 36, 260), (1000, 100, 260)])
@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
def test_tf32_single_pass(sk, M, K, N, a_t, b_t):
    a, b, da, db = operands(sk, M, K, N, a_t, b_t, 7)
    got = sk.asnumpy(sk.matmul(da, db, algo=sk.MM_TF32))
    assert err_ratio(got, a, b) <= 2e-3


@pytest.mark.parametrize("algo", ["x3", "bf16"])
@pytest.mark.parametrize("dist", ["uniform", "positive"])
@pytest.mark.parametrize("M,K,N", [(256, 8192, 256), (512, 16384, 128), (4096, 8192, 512)])
def test_long_k_accumulation_is_promoted(sk, M, K, N, dist, algo):
    """The tensor core accumulates into TMEM with truncation (bias ~ K * 2^-25: 6e-5 at
    K = 8192, measured).  The parity kinds drain TMEM into round-to-nearest fp32 registers
    every few stages; with that the worst case (all-positive data, long K -- the dW GEMM of
    the batch-8192 model has K = 8192) stays within 1e-5 of the exact result."""
    rng = np.random.default_rng(K)
    lo = -1.0 if dist == "uniform" else 0.0
    a = rng.uniform(lo, 1, (M, K)).astype("float32")
    b = rng.uniform(lo, 1, (K, N)).astype("float32")
    if algo == "bf16":
        a, b = bf16_round(a), bf16_round(b)
        got = sk.asnumpy(sk.matmul(sk.to_bf16(sk.array(a)), sk.to_bf16(sk.array(b))))
    else:
        got = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b), algo=sk.MM_TF32X3))
    exact = a.astype(np.float64) @ b.astype(np.float64)
    err = np.abs(got - exact)
    assert err.max() <= 1e-5 * np.abs(exact).max()
    assert np.sqrt((err ** 2).mean()) <= 3e-6 * np.sqrt((exact ** 2).mean())


def bf16_round(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("M,K,N", [(256, 384, 512), (136, 40, 264), (1024, 1024, 1024), (64, 64, 64)])
@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
def test_bf16(sk, M, K, N, a_t, b_t):
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, (M, K)).astype("float32")
    b = rng.uniform(-1, 1, (K, N)).astype("float32")
    ar, br = bf16_round(a), bf16_round(b)
    da = sk.to_bf16(sk.array(np.ascontiguousarray(a.T))).T if a_t else sk.to_bf16(sk.array(a))
    db = sk.to_bf16(sk.array(np.ascontiguousarray(b.T))).T if b_t else sk.to_bf16(sk.array(b))
    assert np.array_equal(sk.asnumpy(da), ar) and np.array_equal(sk.asnumpy(db), br)   # RNE cast is bit-exact
    got = sk.asnumpy(sk.matmul(da, db))
    assert err_ratio(got, ar, br) <= 1e-5


def test_linear_epilogues_tc(sk):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((512, 784)).astype("float32")
    w = (rng.standard_normal((784, 256)) * 0.05).astype("float32")
    bias = rng.standard_normal(256).astype("float32")
    pre = (x.astype(np.float64) @ w.astype(np.float64) + bias).astype("float32")
    for relu in (False, True):
        got = sk.asnumpy(sk.linear(sk.array(x), sk.array(w), sk.array(bias), relu=relu, algo=sk.MM_TF32X3))
        want = np.maximum(pre, 0) if relu else pre
        assert np.abs(got - want).max() <= 1e-5 * np.abs(pre).max()


def test_auto_dispatch_and_views(sk):
    """AUTO uses tcgen05 for large problems and the FFMA kernel for small / skinny ones;
    row-sliced (pitched) operands are consumed in place."""
    rng = np.random.default_rng(5)
    big = rng.uniform(-1, 1, (600, 520)).astype("float32")
    d = sk.array(big)
    a, da = big[40:552, :512], d[40:552, :512]            # pitched K-major view
    b, db = big[:512, 8:264], d[:512, 8:264]              # pitched N-major view
    sk.profile_reset(); sk.profile_enable(True)
    got = sk.asnumpy(sk.matmul(da, db))
    fam = sk.profile_collect()
    sk.profile_enable(False)
    assert "gemm_tc" in fam and "gemm_simt" not in fam
    assert err_ratio(got, a, b) <= 1e-5
    sk.profile_reset(); sk.profile_enable(True)
    small = sk.asnumpy(sk.matmul(sk.array(big[:100, :100]), sk.array(big[:100, :10])))
    fam = sk.profile_collect()
    sk.profile_enable(False)
    assert "gemm_simt" in fam and "gemm_tc" not in fam
    assert err_ratio(small, big[:100, :100], big[:100, :10]) <= 1e-5


@pytest.mark.parametrize("M,K,N", [(8192, 4096, 10), (4096, 8192, 10), (8192, 10, 4096), (1000, 514, 130)])
def test_unaligned_pitch_is_repitched_onto_tcgen05(sk, M, K, N):
    """The 10-class output layer (weight pitch 40 B) and other pitches TMA cannot describe
    are copied into a 16-byte-pitched scratch and still run on the tensor cores."""
    rng = np.random.default_rng(M + N)
    a = rng.uniform(-1, 1, (M, K)).astype("float32")
    b = rng.uniform(-1, 1, (K, N)).astype("float32")
    sk.profile_reset(); sk.profile_enable(True)
    got = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b)))
    got_t = sk.asnumpy(sk.matmul(sk.array(np.ascontiguousarray(a.T)).T, sk.array(b)))
    fam = sk.profile_collect()
    sk.profile_enable(False)
    assert "gemm_tc" in fam and "gemm_simt" not in fam
    assert err_ratio(got, a, b) <= 2e-6 and err_ratio(got_t, a, b) <= 2e-6


# ------------------------------------------------------------------ fp16x3 (the default fp32 path)
F16_SHAPES = [(256, 64, 128), (256, 384, 512), (512, 784, 256), (1000, 100, 260), (260, 72, 136),
              (2048, 1024, 512), (384, 8192, 256), (4096, 200, 4096)]


@pytest.mark.parametrize("M,K,N", F16_SHAPES)
@pytest.mark.parametrize("a_t", [False, True])
@pytest.mark.parametrize("b_t", [False, True])
def test_f16x3_all_layouts(sk, M, K, N, a_t, b_t):
    """fp16 hi/lo splits with one power-of-two scale per row of A / column of B, three
    kind::f16 MMAs per K step: every product is exact in fp32, the dropped lo*lo term is
    2^-24 relative -- tighter than 3xTF32 at twice its rate."""
    a, b, da, db = operands(sk, M, K, N, a_t, b_t, M + K + N)
    got = sk.asnumpy(sk.matmul(da, db, algo=sk.MM_F16X3))
    assert got.shape == (M, N) and got.dtype == np.float32
    assert err_ratio(got, a, b) <= 1e-6


@pytest.mark.parametrize("dist", ["rows_1e-20..1e20", "lognormal", "positive", "sparse", "constant"])
@pytest.mark.parametrize("a_t,b_t", [(False, False), (True, False), (False, True)])
def test_f16x3_scaling_handles_dynamic_range(sk, dist, a_t, b_t):
    """Row / column magnitudes spanning the whole fp32 range, heavy-tailed entries, exact
    zeros and constant matrices: the K-invariant scales keep the 1e-5 bound per element."""
    M, K, N = 512, 640, 384
    rng = np.random.default_rng(42)
    a = rng.uniform(-1, 1, (M, K))
    b = rng.uniform(-1, 1, (K, N))
    if dist == "rows_1e-20..1e20":
        a *= 10.0 ** rng.uniform(-17, 17, (M, 1))
        b *= 10.0 ** rng.uniform(-17, 17, (1, N))
    elif dist == "lognormal":
        a *= np.exp(rng.normal(0, 2.0, (M, K)))
        b *= np.exp(rng.normal(0, 2.0, (K, N)))
    elif dist == "positive":
        a, b = np.abs(a), np.abs(b)
    elif dist == "sparse":
        a *= rng.random((M, K)) < 0.05
        b *= rng.random((K, N)) < 0.05
        a[7] = 0.0
        b[:, 11] = 0.0
    elif dist == "constant":
        a[:], b[:] = 0.1, 0.3
    a, b = a.astype("float32"), b.astype("float32")
    da = sk.array(np.ascontiguousarray(a.T)).T if a_t else sk.array(a)
    db = sk.array(np.ascontiguousarray(b.T)).T if b_t else sk.array(b)
    got = sk.asnumpy(sk.matmul(da, db, algo=sk.MM_F16X3))
    assert np.isfinite(got).all()
    assert err_ratio(got, a, b) <= 4e-6


def test_f16x3_epilogues_and_auto(sk):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((512, 784)).astype("float32")
    w = (rng.standard_normal((784, 256)) * 0.05).astype("float32")
    bias = rng.standard_normal(256).astype("float32")
    pre = (x.astype(np.float64) @ w.astype(np.float64) + bias).astype("float32")
    for relu in (False, True):
        for algo in (sk.MM_F16X3, None):    # None = AUTO, which picks fp16x3 for this shape
            got = sk.asnumpy(sk.linear(sk.array(x), sk.array(w), sk.array(bias), relu=relu, algo=algo))
            want = np.maximum(pre, 0) if relu else pre
            assert np.abs(got - want).max() <= 2e-6 * np.abs(pre).max()
    with pytest.raises(RuntimeError):       # too small for the CTA-pair kernel: explicit request fails loudly
        sk.matmul(sk.array(x[:64]), sk.array(w), algo=sk.MM_F16X3)


@pytest.mark.parametrize("dist", ["uniform", "positive"])
@pytest.mark.parametrize("M,K,N", [(256, 8192, 256), (512, 16384, 128), (4096, 8192, 512)])
def test_f16x3_long_k(sk, M, K, N, dist):
    rng = np.random.default_rng(K)
    lo = -1.0 if dist == "uniform" else 0.0
    a = rng.uniform(lo, 1, (M, K)).astype("float32")
    b = rng.uniform(lo, 1, (K, N)).astype("float32")
    got = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b), algo=sk.MM_F16X3))
    exact = a.astype(np.float64) @ b.astype(np.float64)
    err = np.abs(got - exact)
    assert err.max() <= 1e-5 * np.abs(exact).max()
    assert np.sqrt((err ** 2).mean()) <= 3e-6 * np.sqrt((exact ** 2).mean())


@pytest.mark.parametrize("Bn,I,O", [(512, 384, 256), (1024, 784, 512), (300, 260, 136), (100, 784, 100), (256, 64, 128)])
@pytest.mark.parametrize("dist", ["uniform", "sample_scales"])
def test_linear_bwd_shared_split(sk, Bn, I, O, dist):
    """sk_linear_bwd: dX = adj @ W.T and dW = X.T @ adj from ONE row split of adj; in the dW
    GEMM adj's per-sample scale is folded into X (exact powers of two).  Per-sample gradient
    magnitudes spanning 1e-12..1e6 must not cost accuracy.  Small shapes take the two-GEMM
    fallback inside the same entry point."""
    rng = np.random.default_rng(Bn + I + O)
    adj = rng.uniform(-1, 1, (Bn, O))
    x = rng.uniform(-1, 1, (Bn, I))
    w = rng.uniform(-1, 1, (I, O))
    if dist == "sample_scales":
        adj *= 10.0 ** rng.uniform(-12, 6, (Bn, 1))
        x *= 10.0 ** rng.uniform(-3, 3, (1, I))
    adj, x, w = adj.astype("float32"), x.astype("float32"), w.astype("float32")
    dx, dw = sk.linear_bwd(sk.array(adj), sk.array(x), sk.array(w))
    dx, dw = sk.asnumpy(dx), sk.asnumpy(dw)
    assert dx.shape == (Bn, I) and dw.shape == (I, O)
    assert err_ratio(dx, adj, np.ascontiguousarray(w.T)) <= 2e-6
    assert err_ratio(dw, np.ascontiguousarray(x.T), adj) <= 2e-6


@pytest.mark.parametrize("B_,I,O", [(8192, 4096, 4096), (1024, 512, 1024), (300, 256, 512), (50, 20, 30), (4096, 4096, 8)])
def test_linear_bwd_bias_gradient_from_the_adj_split(sk, B_, I, O):
    """sk_linear_bwd_bias: db = adj.sum(0) (autodiff.pyx:84) as a by-product of the row split of adj
    on the fp16x3 path (and from sk_colsum on the general path); dX / dW unchanged by it."""
    rng = np.random.default_rng(B_ + I + O)
    adj = rng.standard_normal((B_, O)).astype("float32") * np.exp(rng.standard_normal((B_, 1))).astype("float32")
    x = rng.standard_normal((B_, I)).astype("float32")
    w = (rng.standard_normal((I, O)) / np.sqrt(I)).astype("float32")
    d_adj, d_x, d_w = sk.array(adj), sk.array(x), sk.array(w)
    dx, dw, db = sk.linear_bwd(d_adj, d_x, d_w, True)
    dx0, dw0 = sk.linear_bwd(d_adj, d_x, d_w)
    assert np.array_equal(sk.asnumpy(dx), sk.asnumpy(dx0)) and np.array_equal(sk.asnumpy(dw), sk.asnumpy(dw0))
    got = sk.asnumpy(db)
    assert got.shape == (O,) and got.dtype == np.float32
    want = adj.astype(np.float64).sum(0)
    bound = np.abs(adj).astype(np.float64).sum(0)
    assert np.all(np.abs(got - want) <= 1e-6 * bound + 1e-30)       # fp32 tree sum against float64
    assert np.abs(got - adj.sum(0)).max() <= 1e-5 * np.abs(want).max()
