"""BASELINE.json configs[0]: examples/mlp_resnet MLPResNet on MNIST-shaped synthetic data
(784 -> hidden 100, 3 residual blocks, LayerNorm, batch 100, SGD lr 0.01, dropout 0.01),
the epoch loop of examples/mlp_resnet/model.py:72-95 (loss.item() every step), on the
sm_100a backend and on the reference's CPU backend (oracle/_ref) on the host cores.

    python scripts/c1_bench.py [--steps 600] [--variant fn|verbatim]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=600)
ap.add_argument("--warmup", type=int, default=20)
ap.add_argument("--variant", default="fn", choices=["fn", "verbatim"])
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--graph", action="store_true", help="also time the step as one CUDA-graph launch (soket_b200.graph.StaticStep)")
ap.add_argument("--loader", action="store_true",
                help="also time one EPOCH (60 000 samples, 600 steps) fed by the DataLoader: resident gather path, "
                     "the reference's per-sample path on the GPU device, and the reference's own loader on the CPU")
args = ap.parse_args()
DIM, HID, NB, C, B = 784, 100, 3, 10, 100
rng = np.random.default_rng(0)
X = rng.random((B * 64, DIM), dtype=np.float32)        # 64 distinct batches, cycled
y = rng.integers(0, C, B * 64).astype(np.uint8)

from oracle import ref_model                            # noqa: E402  (model builder shared with the tests)


def run(nn, Tensor, SGD, kaiming, sync, to_dev):
    np.random.seed(0)
    model = ref_model.build_model(nn, DIM, HID, NB, C, norm="layer", drop_prob=0.01, retain_fn=args.variant == "fn")
    for m in model.modules():
        if type(m).__name__ == "Linear":
            kaiming(m.weight)
    opt = SGD(model.parameters(), lr=0.01)
    crit = nn.SoftmaxCrossEntropyLoss()
    model.train(True)
    batches = [(to_dev(X[i * B:(i + 1) * B]), to_dev(y[i * B:(i + 1) * B])) for i in range(64)]

    def step(i):
        xb, yb = batches[i % 64]
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss.item()
    for i in range(args.warmup):
        step(i)
    sync()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = step(i)
    sync()
    return (time.perf_counter() - t0) / args.steps, last


out = {"config": f"MLPResNet(784, 100, 3 blocks, LayerNorm, dropout 0.01), batch 100, SGD lr=0.01, variant={args.variant}",
       "steps": args.steps}
import soket_b200 as sk                                 # noqa: E402
import soket_b200.api as soket                          # noqa: E402
from soket_b200 import nn as gnn                        # noqa: E402
from soket_b200.optim import SGD as GSGD                # noqa: E402
sk.init(0)
n0 = sk.launch_count()
sec, loss = run(gnn, soket.Tensor, GSGD, gnn.kaiming_normal, sk.synchronize, lambda a: soket.Tensor(a))
launches = (sk.launch_count() - n0) / (args.steps + args.warmup)
out["gpu"] = {"ms_per_step": sec * 1e3, "samples_per_s": B / sec, "launches_per_step": launches, "last_loss": loss}
if args.graph:
    from soket_b200.graph import StaticStep
    np.random.seed(0)
    model = ref_model.build_model(gnn, DIM, HID, NB, C, norm="layer", drop_prob=0.01, retain_fn=args.variant == "fn")
    for m in model.modules():
        if type(m).__name__ == "Linear":
            gnn.kaiming_normal(m.weight)
    opt = GSGD(model.parameters(), lr=0.01)
    crit = gnn.SoftmaxCrossEntropyLoss()
    model.train(True)
    dev_batches = [(sk.array(X[i * B:(i + 1) * B]), sk.array(y[i * B:(i + 1) * B])) for i in range(64)]
    xb, yb = soket.Tensor(X[:B]), soket.Tensor(y[:B])

    def gstep():
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss
    g = StaticStep(gstep)

    def run_graph(i):
        bx, by = dev_batches[i % 64]
        xb._data[:] = bx               # device-side copy of the batch into the static input buffers
        yb._data[:] = by
        g.launch()
        return g.loss.item()           # loss.item() every step, as examples/mlp_resnet/model.py:89
    for i in range(args.warmup):
        run_graph(i)
    sk.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = run_graph(i)
    sk.synchronize()
    sec_g = (time.perf_counter() - t0) / args.steps
    out["gpu_graph"] = {"ms_per_step": sec_g * 1e3, "samples_per_s": B / sec_g, "last_loss": last,
                        "launches_per_step": "1 graph + 2 batch copies"}
    g.close()


def epoch_through_loader(nn, SGD, kaiming, sync, loader, graph=False):
    """examples/mlp_resnet/model.py:72-97 (mlp_resnet_epoch, training branch) without the
    accuracy pass: for X, y in loader: loss = crit(model(X), y); backward; step; loss.item()."""
    np.random.seed(0)
    model = ref_model.build_model(nn, DIM, HID, NB, C, norm="layer", drop_prob=0.01, retain_fn=args.variant == "fn")
    for m in model.modules():
        if type(m).__name__ == "Linear":
            kaiming(m.weight)
    opt = SGD(model.parameters(), lr=0.01)
    crit = nn.SoftmaxCrossEntropyLoss()
    model.train(True)
    g = None
    if graph:
        from soket_b200.graph import StaticStep
        xb, yb = soket.Tensor(X[:B]), soket.Tensor(y[:B])

        def gstep():
            loss = crit(model(xb), yb)
            loss.backward()
            opt.step()
            return loss
        g = StaticStep(gstep)
    total = 0.0
    for ep in range(2):                       # epoch 0 warms up (allocator, dataset upload); epoch 1 is timed
        sync()
        t0 = time.perf_counter()
        total = 0.0
        for bx, by in loader:
            if g is not None:
                xb._data[:] = bx._data
                yb._data[:] = by._data
                g.launch()
                total += g.loss.item()
            else:
                loss = crit(model(bx), by)
                loss.backward()
                opt.step()
                total += loss.item()
        sync()
        sec = time.perf_counter() - t0
    if g is not None:
        g.close()
    return {"epoch_s": sec, "samples_per_s": loader.max_iter * B / sec, "steps": loader.max_iter,
            "mean_loss": total / loader.max_iter}


if args.loader:
    from soket_b200.transforms import ToTensor as GToTensor
    from soket_b200.utils.data import ArrayDataset, DataLoader as GLoader
    N = 60000
    XL = rng.random((N, DIM), dtype=np.float32)
    yL = rng.integers(0, C, N).astype(np.uint8)
    ds = ArrayDataset(XL, yL, GToTensor(), GToTensor())

    def arm(name, *a, **kw):
        try:
            out[name] = epoch_through_loader(*a, **kw)
        except Exception as e:                # keep the other arms' numbers
            out[name] = {"error": f"{type(e).__name__}: {e}"}
    arm("epoch_gpu_resident_loader",
        gnn, GSGD, gnn.kaiming_normal, sk.synchronize, GLoader(ds, batch_size=B, shuffle=True))
    arm("epoch_gpu_resident_loader_graph",
        gnn, GSGD, gnn.kaiming_normal, sk.synchronize, GLoader(ds, batch_size=B, shuffle=True), graph=True)
    arm("epoch_gpu_per_sample_loader",
        gnn, GSGD, gnn.kaiming_normal, sk.synchronize, GLoader(ds, batch_size=B, shuffle=True, resident=False))
    if not args.no_cpu and ref_model.import_reference() is not None:
        import soket.nn as rnn0
        from soket.nn.init import kaiming_normal as rkaiming
        from soket.optim import SGD as RSGD0
        from soket.transforms import ToTensor as RToTensor
        from soket.utils.data import DataLoader as RLoader

        class RefArrays(ArrayDataset):        # same samples, the reference's ToTensor (CPU tensors)
            pass
        rds = RefArrays(XL, yL, RToTensor(), RToTensor())
        arm("epoch_cpu_reference_loader",
            rnn0, RSGD0, rkaiming, lambda: None, RLoader(rds, batch_size=B, shuffle=True))
        out["epoch_cpu_reference_loader"]["cores"] = os.cpu_count()

if not args.no_cpu:
    ref = ref_model.import_reference()
    if ref is not None:
        import soket.nn as rnn
        from soket.nn.init import kaiming_normal
        from soket.optim import SGD as RSGD
        best = None
        sec_c, loss_c = run(rnn, ref.Tensor, RSGD, kaiming_normal, lambda: None, lambda a: ref.Tensor(a))
        out["cpu_reference"] = {"ms_per_step": sec_c * 1e3, "samples_per_s": B / sec_c, "cores": os.cpu_count(),
                                "last_loss": loss_c, "openblas_threads": os.environ.get("OPENBLAS_NUM_THREADS", "default")}
        out["speedup"] = sec_c / sec
print(json.dumps(out), flush=True)
