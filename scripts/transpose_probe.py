import sys
sys.path.insert(0, '.')
import numpy as np
import soket_b200 as sk
sk.init(0)
shapes = [(65536, 4096), (4096, 65536), (16384, 16384), (65000, 4000), (8192, 4096), (4096, 8192), (60000, 4100)]
if len(sys.argv) > 1:
    shapes = shapes[:1]
for (R, C) in shapes:
    a = sk.random.uniform(0, 1, (R, C), dtype='float32')
    fn = lambda: sk.ascontiguousarray(a.T)
    for _ in range(3): fn()
    ts = []
    for _ in range(1 if len(sys.argv) > 1 else 8):
        sk.flush_l2()
        e0, e1 = sk.Event(), sk.Event(); e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_ms(e1))
    ms = float(np.median(ts))
    print(f'transpose ({R},{C}) -> ({C},{R}): {ms:.3f} ms  {8.0 * R * C / ms / 1e6:.0f} GB/s', flush=True)
    del a
    sk.empty_cache()
