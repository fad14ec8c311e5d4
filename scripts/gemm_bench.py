import sys, os
sys.path.insert(0, '.')
import soket_b200 as sk
sk.init(0)
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
names = {sk.MM_TF32X3: 'x3', sk.MM_TF32: 'tf32', sk.MM_BF16: 'bf16'}
algos = [int(x) for x in os.environ.get('ALGOS', f'{sk.MM_TF32X3}').split(',')]
for (M, K, N, at, bt) in [(8192, 4096, 4096, False, False), (8192, 4096, 4096, False, True), (4096, 8192, 4096, True, False), (8192, 784, 4096, False, False)]:
    for algo in algos:
        ms, tf = bench(M, K, N, algo, at, bt)
        print(f'  {names[algo]:5s} M{M} K{K} N{N} aT={at} bT={bt}: {ms:8.3f} ms  {tf:8.1f} TFLOP/s')
