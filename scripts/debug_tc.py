import sys, numpy as np
sys.path.insert(0, '.')
import soket_b200 as sk
sk.init(0)
np.set_printoptions(linewidth=200, precision=4, suppress=True)
def run(name, a, b, algo, bf16=False):
    try:
        da, db = sk.array(a), sk.array(b)
        if bf16: da, db = sk.to_bf16(da), sk.to_bf16(db)
        out = sk.asnumpy(sk.matmul(da, db, algo=algo))
        ref = a.astype(np.float64) @ b.astype(np.float64)
        print(f'{name}: out[0,:6]={out[0,:6]} ref[0,:6]={ref[0,:6]} max|out|={np.abs(out).max():.4f} maxerr={np.abs(out-ref).max():.4e} nz={np.count_nonzero(out)}')
        return out
    except Exception as e:
        print(f'{name}: EXC {e}')
M = K = N = 128
ones_a = np.ones((M, K), 'float32'); ones_b = np.ones((K, N), 'float32')
for algo, nm, bf in ((sk.MM_TF32, 'tf32', False), (sk.MM_BF16, 'bf16', True), (sk.MM_TF32X3, 'x3', False)):
    run(nm + ' ones', ones_a, ones_b, algo, bf)
rng = np.random.default_rng(0)
a = rng.integers(-3, 4, (M, K)).astype('float32'); b = rng.integers(-3, 4, (K, N)).astype('float32')
for algo, nm, bf in ((sk.MM_TF32, 'tf32', False), (sk.MM_BF16, 'bf16', True), (sk.MM_TF32X3, 'x3', False)):
    o = run(nm + ' ints', a, b, algo, bf)
# row/col structure probes: A = row index, B = identity
ar = np.tile(np.arange(M, dtype='float32')[:, None], (1, K)); eye = np.eye(K, N, dtype='float32')
o = run('tf32 rowidx@eye', ar, eye, sk.MM_TF32)
if o is not None: print(' col0', o[:8, 0], ' diag', np.diag(o)[:8])
ak = np.tile(np.arange(K, dtype='float32')[None, :], (M, 1))
o = run('tf32 kidx@eye', ak, eye, sk.MM_TF32)
if o is not None: print(' row0', o[0, :40])
