"""LayerNorm forward/backward kernels at the wide-model shape: algorithmic GB/s (CUDA events, L2 flushed)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import soket_b200 as sk
from soket_b200 import _fused as F
sk.init(0)
B, H = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = sk.random.uniform(-1, 1, (B, H), dtype='float32')
adj = sk.random.uniform(-1, 1, (B, H), dtype='float32')
g = sk.ones((H,), 'float32'); b = sk.zeros((H,), 'float32')
ln, mean, rstd = F.layernorm_fwd(x, g, b, None, 1e-5, True)
ln2, mean2, rstd2 = F.layernorm_fwd(x, g, b, adj, 1e-5, True)


def t(fn, nbytes, name):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        sk.flush_l2()
        e0, e1 = sk.Event(), sk.Event(); e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_ms(e1))
    ms = float(np.median(ts))
    print(f'{name:34s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s', flush=True)


n = B * H
t(lambda: F.layernorm_fwd(x, g, b, None, 1e-5, True), 8 * n, 'ln_fwd (+relu)')
t(lambda: F.layernorm_fwd(x, g, b, adj, 1e-5, True), 12 * n, 'ln_fwd (+residual +relu)')
t(lambda: F.layernorm_bwd(adj, x, g, b, mean, rstd, None, 1), 12 * n, 'ln_bwd (mask recomputed)')
t(lambda: F.layernorm_bwd(adj, x, g, b, mean2, rstd2, ln2, 2, True), 20 * n, 'ln_bwd (mask from y, +dresidual)')
