"""One launch of each hot kernel at the wide-model shapes (B = 8192, H = 4096), for
`ncu --set full --clock-control none --import-source on -o gpurun_out/prof python scripts/profile_kernels.py`."""
import sys
sys.path.insert(0, '.')
import soket_b200 as sk
from soket_b200 import _fused as F
sk.init(0)
B, H = 8192, 4096
x = sk.random.uniform(-1, 1, (B, H), dtype='float32')
w = sk.random.uniform(-1, 1, (H, H), dtype='float32')
adj = sk.random.uniform(-1, 1, (B, H), dtype='float32')
g = sk.ones((H,), 'float32'); b = sk.zeros((H,), 'float32')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for rep in range(reps):
    y = sk.linear(x, w, b, relu=True)                         # fwd: fp16x3, K-major A, N-major B, bias+relu epilogue
    dx = sk.matmul(adj, w.T)                                  # dX: K-major B
    dw = sk.matmul(x.T, adj)                                  # dW: M-major A, N-major B, K = 8192
    y3 = sk.linear(x, w, b, relu=True, algo=sk.MM_TF32X3)     # the 3xTF32 alternative
    yb = sk.matmul(sk.to_bf16(x), sk.to_bf16(w))              # bf16 sweep kernel
    yt = sk.matmul(x, w, algo=sk.MM_TF32)                     # single-pass tf32
    ln, mean, rstd = F.layernorm_fwd(x, g, b, None, 1e-5, True)
    dln = F.layernorm_bwd(adj, x, g, b, mean, rstd, None, 1)
    ln2, mean2, rstd2 = F.layernorm_fwd(x, g, b, adj, 1e-5, True)
    dln2 = F.layernorm_bwd(adj, x, g, b, mean2, rstd2, ln2, 2, True)
    s = sk.add(x, adj)
    r = sk.relu_backward(x, adj)
    cs = F.colsum(adj)
    rs = sk.sum(x, (1,), 'float32', None, True)
    fs = sk.sum(x, None, 'float32', None, False)
    mx = sk.max(x, (1,), None, True)
    t = sk.ascontiguousarray(x.T)
    m = [sk.zeros((H, H), 'float32') for _ in range(2)]
    F.adam_step([w], [dw], [m[0]], [m[1]], 1e-3, 0.9, 0.999, 1e-8, 0.0, 0.1, 0.001, rep == 0, 1.0)
    sk.synchronize()
print('done')
