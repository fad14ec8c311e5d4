import sys, time, numpy as np
sys.path.insert(0, '.')
import soket_b200 as sk
import bench
from soket_b200 import nn
from soket_b200.optim import Adam
import soket_b200.api as soket
sk.init(0)
sk.random.seed(1234)
model = bench.build_model(nn, 4096, 8)
for m in model.modules():
    if type(m).__name__ == 'Linear': nn.kaiming_normal(m.weight)
opt = Adam(list(model.parameters()), lr=1e-3)
crit = nn.SoftmaxCrossEntropyLoss()
Xh, yh = bench.synthetic_batch(8192, 100)
Xd, yd = soket.Tensor(Xh), soket.Tensor(yh)
for i in range(14):
    sk.synchronize(); t0 = time.perf_counter(); e0, e1 = sk.Event(), sk.Event(); e0.record()
    loss = crit(model(Xd), yd); loss.backward(); opt.step()
    e1.record(); t1 = time.perf_counter(); e1.synchronize(); t2 = time.perf_counter()
    st = sk.memory_stats()
    print(f'step {i}: device {e0.elapsed_ms(e1):8.2f} ms  host-enqueue {(t1-t0)*1e3:8.2f} ms  wall {(t2-t0)*1e3:8.2f} ms  reserved {st["reserved"]/2**30:.2f} GiB in_use {st["in_use"]/2**30:.2f} peak {st["peak_in_use"]/2**30:.2f}')
