import sys, time, numpy as np
sys.path.insert(0, '.')
import soket_b200 as sk
sk.init(0)
rng = np.random.default_rng(0)
def ratio(got, a, b):
    exact = a.astype(np.float64) @ b.astype(np.float64)
    bound = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    e = np.abs(got.astype(np.float64) - exact)
    return (e / bound).max(), (e.max() / np.abs(exact).max()), np.sqrt((e**2).mean()) / np.sqrt((exact**2).mean())
for (M, K, N) in [(256, 256, 256), (256, 1024, 256), (256, 4096, 256), (256, 8192, 256), (256, 16384, 256), (512, 8192, 512)]:
    for dist in ('uniform', 'positive'):
        a = rng.uniform(-1, 1, (M, K)).astype('float32') if dist == 'uniform' else rng.uniform(0, 1, (M, K)).astype('float32')
        b = rng.uniform(-1, 1, (K, N)).astype('float32') if dist == 'uniform' else rng.uniform(0, 1, (K, N)).astype('float32')
        da, db = sk.array(a), sk.array(b)
        out = {}
        for nm, algo in (('x3', sk.MM_TF32X3), ('simt', sk.MM_SIMT), ('tf32', sk.MM_TF32)):
            out[nm] = ratio(sk.asnumpy(sk.matmul(da, db, algo=algo)), a, b)
        out['numpy'] = ratio(np.matmul(a, b), a, b)
        print(f'M{M} K{K} N{N} {dist:8s} ' + '  '.join(f'{k}: bound-rel {v[0]:.2e} max-rel {v[1]:.2e} rms-rel {v[2]:.2e}' for k, v in out.items()))
# timing of the main shapes
def bench(M, K, N, algo, a_t=False, b_t=False, reps=10):
    a = sk.random.uniform(-1, 1, (K, M) if a_t else (M, K), dtype='float32'); b = sk.random.uniform(-1, 1, (N, K) if b_t else (K, N), dtype='float32')
    if a_t: a = a.T
    if b_t: b = b.T
    if algo == sk.MM_BF16: a, b = sk.to_bf16(sk.ascontiguousarray(a)), sk.to_bf16(sk.ascontiguousarray(b))
    for _ in range(3): sk.matmul(a, b, algo=algo)
    e0, e1 = sk.Event(), sk.Event(); e0.record()
    for _ in range(reps): sk.matmul(a, b, algo=algo)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_ms(e1) / reps
    return ms, 2.0 * M * N * K / ms / 1e9
for (M, K, N, at, bt) in [(8192, 4096, 4096, False, False), (8192, 4096, 4096, False, True), (4096, 8192, 4096, True, False), (8192, 784, 4096, False, False), (4096, 4096, 4096, False, False), (8192, 8192, 8192, False, False)]:
    for nm, algo in (('x3', sk.MM_TF32X3), ('tf32', sk.MM_TF32), ('bf16', sk.MM_BF16)):
        ms, tf = bench(M, K, N, algo, at, bt)
        print(f'  {nm:5s} M{M} K{K} N{N} aT={at} bT={bt}: {ms:8.3f} ms  {tf:8.1f} TFLOP/s')
