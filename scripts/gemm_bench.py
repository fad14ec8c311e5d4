"""GEMM sweep on one B200: TFLOP/s (2*M*N*K, CUDA events) and error vs float64 per algorithm.
    ALGOS=5,2,3,4 python scripts/gemm_bench.py [--sweep]"""
import os
import sys
import numpy as np
sys.path.insert(0, '.')
import soket_b200 as sk
sk.init(0)


def bench(M, K, N, algo, a_t=False, b_t=False, reps=10):
    a = sk.random.uniform(-1, 1, (K, M) if a_t else (M, K), dtype='float32')
    b = sk.random.uniform(-1, 1, (N, K) if b_t else (K, N), dtype='float32')
    if a_t: a = a.T
    if b_t: b = b.T
    if algo == sk.MM_BF16:
        a, b = sk.to_bf16(sk.ascontiguousarray(a)), sk.to_bf16(sk.ascontiguousarray(b))
    for _ in range(3):
        sk.matmul(a, b, algo=algo)
    e0, e1 = sk.Event(), sk.Event()
    e0.record()
    for _ in range(reps):
        sk.matmul(a, b, algo=algo)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / reps
    return ms, 2.0 * M * N * K / ms / 1e9


def error(M, K, N, algo, dist):
    rng = np.random.default_rng(1)
    lo = 0.0 if dist == 'positive' else -1.0
    a = rng.uniform(lo, 1, (M, K)).astype('float32'); b = rng.uniform(lo, 1, (K, N)).astype('float32')
    got = sk.asnumpy(sk.matmul(sk.array(a), sk.array(b), algo=algo)).astype(np.float64)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    bound = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    e = np.abs(got - exact) / bound
    return e.max(), np.sqrt((e ** 2).mean())


names = {sk.MM_TF32X3: 'tf32x3', sk.MM_TF32: 'tf32', sk.MM_BF16: 'bf16', sk.MM_F16X3: 'f16x3', sk.MM_SIMT: 'simt'}
algos = [int(x) for x in os.environ.get('ALGOS', f'{sk.MM_F16X3},{sk.MM_TF32X3}').split(',')]
shapes = [(8192, 4096, 4096, False, False), (8192, 4096, 4096, False, True), (4096, 8192, 4096, True, False),
          (8192, 784, 4096, False, False)]
if '--sweep' in sys.argv:
    shapes = [(n, n, n, False, False) for n in (256, 512, 1024, 2048, 4096, 8192, 16384)]
for (M, K, N, at, bt) in shapes:
    for algo in algos:
        ms, tf = bench(M, K, N, algo, at, bt)
        print(f'  {names[algo]:6s} M{M} K{K} N{N} aT={at} bT={bt}: {ms:8.3f} ms  {tf:8.1f} TFLOP/s', flush=True)
if '--errors' in sys.argv:
    for algo in algos:
        if algo in (sk.MM_BF16, sk.MM_TF32):
            continue
        for dist in ('uniform', 'positive'):
            for K in (512, 4096, 16384):
                mx, rms = error(512, K, 256, algo, dist)
                print(f'  err {names[algo]:6s} {dist:8s} K={K:5d}: max {mx:.2e} rms {rms:.2e} (relative to |a|@|b|)', flush=True)
if '--linear' in sys.argv:
    for (M, K, N) in [(8192, 4096, 4096), (100, 784, 100)]:
        x = sk.random.uniform(-1, 1, (M, K), dtype='float32'); w = sk.random.uniform(-1, 1, (K, N), dtype='float32')
        bias = sk.random.uniform(-1, 1, (N,), dtype='float32')
        for _ in range(3):
            sk.linear(x, w, bias, relu=True)
        e0, e1 = sk.Event(), sk.Event(); e0.record()
        for _ in range(20):
            sk.linear(x, w, bias, relu=True)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_ms(e1) / 20
        print(f'  fused relu(x@w+b) M{M} K{K} N{N}: {ms:8.4f} ms  {2.0 * M * N * K / ms / 1e9:8.2f} TFLOP/s', flush=True)
