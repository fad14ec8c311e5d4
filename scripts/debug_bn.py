import sys, numpy as np
sys.path.insert(0, '.')
import soket_b200 as sk
sk.init(0)
import soket_b200.api as soket
from soket_b200 import nn
from soket_b200.optim import SGD, Adam
from oracle import ref_model, soket_np as O
sys.path.insert(0, 'tests')
from test_engine_gpu import make_pair, rel

for fuse in (True, False):
    nn.set_fusion(fuse)
    dim, hidden, nb, C, B = 784, 100, 3, 10, 100
    om, model, named = make_pair(sk, 'batch', dim, hidden, nb, C)
    rng = np.random.default_rng(1)
    X = rng.random((B, dim), dtype=np.float32); y = rng.integers(0, C, B).astype(np.uint8)
    logits = model(soket.Tensor(X)); loss = nn.SoftmaxCrossEntropyLoss()(logits, soket.Tensor(y)); loss.backward()
    wl = om.forward(X); om.loss(wl, y); G = om.backward()
    print('fuse', fuse, 'logits rel', rel(logits.numpy(), wl))
    for k, t in named.items():
        g = t.grad.numpy()
        print(f'  {k:14s} max|want| {np.abs(G[k]).max():.3e}  max|err| {np.abs(g-G[k]).max():.3e}')
    # trajectory
    for opt in ('sgd', 'adam'):
        om, model, named = make_pair(sk, 'batch', dim, hidden, nb, C)
        names = om.names()
        if opt == 'sgd':
            oo = O.SGD(len(names), lr=0.01); do = SGD(model.parameters(), lr=0.01)
        else:
            oo = O.Adam(len(names), lr=0.001, weight_decay=0.001); do = Adam(model.parameters(), lr=0.001, weight_decay=0.001)
        crit = nn.SoftmaxCrossEntropyLoss()
        rng = np.random.default_rng(2)
        errs = []
        for s in range(12):
            X = rng.random((B, dim), dtype=np.float32); y = rng.integers(0, C, B).astype(np.uint8)
            loss = crit(model(soket.Tensor(X)), soket.Tensor(y)); loss.backward(); do.step()
            l, _ = om.train_step(X, y, oo)
            perr = max(rel(named[k].numpy(), om.params[k]) for k in names)
            worst = max(names, key=lambda k: rel(named[k].numpy(), om.params[k]))
            errs.append((abs(loss.item() - l), perr, worst))
        print(' ', opt, [(f'{a:.1e}', f'{b:.1e}', w) for a, b, w in errs])
