#!/usr/bin/env python
"""bench.py -- MLPResNet training throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this repo (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's own CPU path
    python bench.py --impl dropin --steps K ...                # the UNMODIFIED reference on soket.gpu()
    torchrun --nproc-per-node N ... bench.py --gpus N ...      # one rank per GPU, NCCL

Workload (BASELINE.json configs[3] / [4]): the wide MLPResNet of
examples/mlp_resnet/model.py -- 784 -> 4096, 8 residual blocks (Linear, LayerNorm, ReLU,
Dropout(0.01), Linear, LayerNorm; residual add; ReLU), -> 10 classes -- in the `self.fn`-retaining
variant that makes all 68 tensors (271.9 M parameters) trainable (quirk Q1; `--verbatim-q1` runs
the 4-tensor model as written), GLOBAL batch 8192, Adam(lr=1e-3), fp32, synthetic MNIST-shaped
data, random-init weights (kaiming_normal, quirk Q9).  One step = forward + softmax-CE + backward
+ Adam update.  Data-parallel runs (N > 1) SPLIT the global batch (rank r takes rows
[r B/N, (r+1) B/N): strong scaling, as section 8e defines the configuration) and all-reduce every
gradient over NCCL, bucketed and overlapped with backward; `--scaling weak` keeps 8192 rows per
GPU instead.

One JSON line on stdout (rank 0):
  value      samples/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e        samples/s through the public API with per-step pinned-host -> device copies of the
             batch and a device -> host read of the loss
  roofline   the dominant kernel (the tcgen05 GEMM): achieved algorithmic TFLOP/s (2*M*N*K), measured
             live with CUDA events around every GEMM kernel launch of K steps, against the measured
             bf16 tensor peak of MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch of that
             kernel from the committed ncu --set full capture (profiles/*_kernel_traffic.json)
  cpu_baseline  the reference (oracle/_ref, else the NumPy port) on the host cores, on a bounded
             sample of the same workload (rank 0, N = 1 only)
  parity_check  (N = 1) one step of THIS model on the sample's rows against the reference's own
             step from identical weights: loss, three gradient norms, three parameter norms
  dp_check   (N > 1) a checksum of every parameter after the timed loops, identical on all ranks
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The CPU legs (the reference arm; the cpu_baseline leg of a 1-GPU run) use every host core.
# torchrun exports OMP_NUM_THREADS=1 to its workers, and OpenBLAS reads its thread count when NumPy
# is first imported: decide it here, before that import.
_NCPU = os.cpu_count() or 1
if "reference" in sys.argv or os.environ.get("WORLD_SIZE", "1") == "1":
    os.environ["OPENBLAS_NUM_THREADS"] = str(_NCPU)
    os.environ["OMP_NUM_THREADS"] = str(_NCPU)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line (the driver parses it): keep a private handle to the real
# stdout and point fd 1 at stderr, so that banners printed by libraries at the C level (e.g.
# "NCCL version ..." when NCCL_DEBUG is set in the environment) cannot land in front of it.
_JSON_OUT = sys.stdout       # main() swaps in the private handle; importers keep plain stdout


def _isolate_stdout():
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


DIM, HIDDEN, BLOCKS, CLASSES = 784, 4096, 8, 10
GLOBAL_BATCH = 8192
DROP_P = 0.01
LR = 1e-3
NORM = "layer"
METRIC = "mlpresnet_train_samples_per_s"


def flops_per_step(batch, hidden=HIDDEN, blocks=BLOCKS, dim=DIM, classes=CLASSES):
    """2*M*N*K over every GEMM of forward + backward (no dX for the first layer)."""
    fwd = 2.0 * batch * (dim * hidden + blocks * 2 * hidden * hidden + hidden * classes)
    bwd = 2.0 * batch * (dim * hidden + blocks * 2 * 2 * hidden * hidden + 2 * hidden * classes)
    return fwd + bwd


def read_env():
    """RANK / LOCAL_RANK / WORLD_SIZE as torchrun sets them (no package import: the reference arm
    must not load the product)."""
    class Env:
        rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        master_addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
        master_port = int(os.environ.get("MASTER_PORT", "29500"))
    return Env()


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info() if p.get("user_api") == "blas"] or [1])
    except Exception:
        return None


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # under load = the upper half of the samples (idle samples bracket the region)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- model builders
def build_model(nn, hidden, blocks, drop_p=DROP_P, norm="layer", retain_fn=True):
    """examples/mlp_resnet/model.py:17-58 out of the given `nn` namespace (the reference's soket.nn
    or soket_b200.nn).  retain_fn keeps `self.fn`, which makes the inner layers visible to
    parameters() / modules() / train() (quirk Q1); False is the model exactly as written there
    (4 trainable tensors: the first and last Linear)."""
    class ResidualBlock(nn.Sequential):
        def __init__(self, dim, hid):
            Norm = nn.LayerNorm if norm == "layer" else nn.BatchNorm1d
            fn = nn.Sequential(nn.Linear(dim, hid), Norm(hid), nn.ReLU(), nn.Dropout(p=drop_p),
                               nn.Linear(hid, hid), Norm(hid))
            super().__init__(nn.Residual(fn), nn.ReLU())
            if retain_fn:
                self.fn = fn

    class MLPResNet(nn.Sequential):
        def __init__(self):
            super().__init__(nn.Linear(DIM, hidden), nn.ReLU(),
                             *[ResidualBlock(hidden, hidden) for _ in range(blocks)],
                             nn.Linear(hidden, CLASSES))

    return MLPResNet()


def synthetic_batch(batch, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((batch, DIM), dtype=np.float32)
    y = rng.integers(0, CLASSES, batch).astype(np.uint8)
    return X, y


def load_traffic():
    """DRAM bytes per launch of the dominant GEMM kernel from the committed ncu capture."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernel_traffic.json")))
    if not files:
        return None, None
    d = json.load(open(files[-1]))
    return d.get("gemm_fwd_8192x4096x4096", {}).get("dram_bytes_per_launch"), os.path.basename(files[-1])


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload_config(args, world, per_gpu):
    trainable = (4 + 8 * args.blocks) if not args.verbatim_q1 else 4
    cfg = {
        "workload": f"MLPResNet(784, hidden={args.hidden}, blocks={args.blocks}, classes=10, "
                    f"{'LayerNorm' if args.norm == 'layer' else 'BatchNorm1d'}, "
                    f"dropout={DROP_P}) train step, Adam lr=1e-3, fp32, {trainable} trainable tensors "
                    f"(BASELINE.json configs[3]; configs[4] for N>1)",
        "global_batch": per_gpu * world, "batch_per_gpu": per_gpu,
        "parallelism": f"dp{world}" if world > 1 else "single",
        "l2": "per-step working set (activations + 1.09 GB of weights) >> 126 MB L2; no explicit flush",
    }
    return cfg


# --------------------------------------------------------------------------- CPU reference
class CpuReference:
    """The reference's own CPU implementation of the step (oracle/_ref), else the NumPy port."""

    def __init__(self, args, install_compat):
        from oracle import ref_model
        self.ref_model = ref_model
        self.soket = ref_model.import_reference(install_compat=install_compat)
        self.kind = "reference" if self.soket is not None else "port"
        self.cores = _NCPU
        hidden, blocks = args.hidden, args.blocks
        if self.soket is not None:
            import soket.nn as rnn
            from soket.nn.init import kaiming_normal
            from soket.optim import Adam
            np.random.seed(0)
            self.model = build_model(rnn, hidden, blocks, norm=args.norm, retain_fn=not args.verbatim_q1)
            for m in self.model.modules():
                if type(m).__name__ == "Linear":
                    kaiming_normal(m.weight)
            self.params = list(self.model.parameters())
            self.opt = Adam(self.model.parameters(), lr=LR)
            self.crit = rnn.SoftmaxCrossEntropyLoss()
            self.model.train(True)
        else:
            from oracle import soket_np as O
            self.om = O.MLPResNet(DIM, hidden, blocks, CLASSES, norm="layer")
            self.om.init_kaiming(0)
            self.oopt = O.Adam(len(self.om.names()), lr=LR)

    def set_batch(self, X, y):
        self.X, self.y = X, y
        if self.soket is not None:
            self.Xt, self.yt = self.soket.Tensor(X), self.soket.Tensor(y)

    def step(self):
        if self.soket is None:
            return self.om.train_step(self.X, self.y, self.oopt)[0]
        loss = self.crit(self.model(self.Xt), self.yt)
        loss.backward()
        self.opt.step()
        return loss.item()

    def numpy_params(self):
        return [self.ref_model.to_numpy(self.soket, p) for p in self.params]

    def numpy_grads(self, idx):
        return [self.ref_model.to_numpy(self.soket, self.params[i].grad) for i in idx]

    def time_steps(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / max(steps, 1)


def run_reference(args, env):
    if env.rank != 0:
        return
    sample = args.cpu_sample_batch
    ref = CpuReference(args, install_compat=False)
    assert "soket_b200" not in sys.modules, "the reference arm must not load the product"
    ref.set_batch(*synthetic_batch(sample, 0))
    sec = ref.time_steps(args.steps, args.warmup)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME workload description as this repo's arm at this N (the driver compares them); what the
        # CPU actually steps through per timed step -- a bounded sample of it -- is in cpu_baseline
        "config": workload_config(args, max(args.gpus, 1),
                                  args.global_batch // max(args.gpus, 1) if args.scaling == "strong" else args.global_batch),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": ref.cores, "kind": ref.kind,
                         "blas_threads": blas_threads(), "rows_per_step": sample,
                         "env": {k: os.environ.get(k) for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS")},
                         "sample": (f"{args.steps} Adam step(s) of the same model on the WHOLE {sample}-row batch per step "
                                    f"after {args.warmup} warm-up step(s); NumPy/OpenBLAS on all host cores"
                                    if sample >= args.global_batch else
                                    f"{args.steps} Adam step(s) of the same model on {sample} rows per step after "
                                    f"{args.warmup} warm-up step(s) (the full batch is {args.global_batch} rows: "
                                    f"--cpu-sample-batch {args.global_batch} --steps 3 times it whole; the per-step Adam "
                                    f"update over 272 M parameters does not shrink with the sample, so full-batch CPU "
                                    f"throughput is ~15 % higher than this sample's); NumPy/OpenBLAS on all host cores")},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# --------------------------------------------------------------------------- parity check (N = 1)
def parity_check(args, ref, sk, soket, nn, Adam):
    """One dropout-free step on the CPU sample's rows, from the reference's own initial weights, on
    both sides (the reference's step doubles as the cpu_baseline leg's warm-up): loss, three
    gradient norms and three parameter norms after the Adam update."""
    if ref.soket is None:
        return {"skipped": "oracle/_ref is not built (NumPy port only)"}
    X, y = ref.X, ref.y
    model = build_model(nn, args.hidden, args.blocks, norm=args.norm, retain_fn=not args.verbatim_q1)
    params = list(model.parameters())
    want0 = ref.numpy_params()
    assert len(params) == len(want0)
    for p, w in zip(params, want0):
        assert tuple(p.shape) == tuple(w.shape), (p.shape, w.shape)
        p.data = soket.Tensor(w)
    del want0
    opt = Adam(params, lr=LR)
    crit = nn.SoftmaxCrossEntropyLoss()
    model.train(False)          # Dropout off on both sides (reachable through self.fn, quirk Q1)
    ref.model.train(False)
    loss = crit(model(soket.Tensor(X)), soket.Tensor(y))
    loss.backward()
    pick = sorted({0, len(params) // 2, len(params) - 2})       # first weight, a middle tensor, last weight
    gnorm_dev = [float(np.linalg.norm(params[i].grad.numpy().astype(np.float64))) for i in pick]
    opt.step()
    loss_dev = float(loss.item())
    loss_ref = float(ref.step())
    ref.model.train(True)
    gnorm_ref = [float(np.linalg.norm(g.astype(np.float64))) for g in ref.numpy_grads(pick)]
    pnorm_dev = [float(np.linalg.norm(params[i].numpy().astype(np.float64))) for i in pick]
    pnorm_ref = [float(np.linalg.norm(ref.ref_model.to_numpy(ref.soket, ref.params[i]).astype(np.float64))) for i in pick]
    rel = lambda a, b: abs(a - b) / max(abs(b), 1e-30)
    out = {
        "rows": int(X.shape[0]), "tensors": pick,
        "loss": {"device": loss_dev, "reference": loss_ref, "rel_err": rel(loss_dev, loss_ref)},
        "grad_norm_rel_err": [rel(a, b) for a, b in zip(gnorm_dev, gnorm_ref)],
        "param_norm_rel_err": [rel(a, b) for a, b in zip(pnorm_dev, pnorm_ref)],
    }
    out["ok"] = bool(out["loss"]["rel_err"] <= 1e-5 and max(out["grad_norm_rel_err"]) <= 1e-4
                     and max(out["param_norm_rel_err"]) <= 1e-5)
    return out


# --------------------------------------------------------------------------- our arm
def run_ours(args, env):
    import soket_b200 as sk
    from soket_b200 import dp, nn
    from soket_b200 import engine as E
    from soket_b200.optim import Adam
    import soket_b200.api as soket

    sk.init(env.local_rank)
    rdv = dp.Rendezvous(env) if env.world > 1 else None

    strong = args.scaling == "strong"
    if strong and args.global_batch % env.world:
        raise SystemExit(f"global batch {args.global_batch} does not split over {env.world} ranks")
    batch = args.global_batch // env.world if strong else args.global_batch     # rows per GPU

    sk.random.seed(1234)          # identical initial weights on every rank
    model = build_model(nn, args.hidden, args.blocks, norm=args.norm, retain_fn=not args.verbatim_q1)
    for m in model.modules():
        if type(m).__name__ == "Linear":
            nn.kaiming_normal(m.weight)
    params = list(model.parameters())
    n_params = sum(int(p.size) for p in params)
    opt = Adam(params, lr=LR)
    ddp = dp.DataParallel(opt, rdv)
    ddp.broadcast_parameters(0)
    crit = nn.SoftmaxCrossEntropyLoss()
    model.train(True)
    sk.random.seed(99 + env.rank)  # dropout masks differ per rank

    if strong:      # rank r takes rows [r B/W, (r+1) B/W) of ONE global batch (section 8e)
        Xg, yg = synthetic_batch(args.global_batch, 100)
        rows = dp.shard_rows(args.global_batch, env.rank, env.world)
        Xh, yh = np.ascontiguousarray(Xg[rows]), np.ascontiguousarray(yg[rows])
        del Xg, yg
    else:
        Xh, yh = synthetic_batch(batch, 100 + env.rank)
    Xd, yd = soket.Tensor(Xh), soket.Tensor(yh)
    pin_x = sk.PinnedBuffer(Xh.shape, "float32"); pin_x.array[...] = Xh
    pin_y = sk.PinnedBuffer(yh.shape, "uint8"); pin_y.array[...] = yh
    pin_loss = [sk.PinnedBuffer((1,), "float32") for _ in range(2)]
    loss_ready = [sk.Event() for _ in range(2)]
    # two staging buffers: the batch of step k+1 crosses PCIe on the copy stream while step k computes
    stage_x = [sk.empty(Xh.shape, "float32") for _ in range(2)]
    stage_y = [sk.empty(yh.shape, "uint8") for _ in range(2)]
    e2e_state = {"k": 0, "primed": False}

    def step_resident():
        loss = crit(model(Xd), yd)
        loss.backward()
        ddp.step()               # joins the gradient all-reduces bucket by bucket, then updates
        return loss

    def step_e2e():
        """One step through the public API with its inputs coming from pinned host memory.  Every
        step copies one batch host -> device (here: the NEXT step's, prefetched on the copy
        stream; the first call also copies its own) and reads the loss back to the host."""
        k = e2e_state["k"]
        if not e2e_state["primed"]:
            pin_x.prefetch_to_device(stage_x[k % 2])
            pin_y.prefetch_to_device(stage_y[k % 2])
            e2e_state["primed"] = True
        sk.prefetch_wait()                                   # this step's batch has landed
        pin_x.prefetch_to_device(stage_x[(k + 1) % 2])       # next step's batch, overlapping this step
        pin_y.prefetch_to_device(stage_y[(k + 1) % 2])
        loss = crit(model(E.Tensor._const(stage_x[k % 2])), E.Tensor._const(stage_y[k % 2]))
        loss.backward()
        ddp.step()
        # the loss of EVERY step is read back to the host; the read of step k is consumed while
        # step k+1 is being enqueued (the host runs one step ahead instead of draining the GPU)
        pin_loss[k % 2].copy_from_device(loss._data.reshape(1))
        loss_ready[k % 2].record()
        e2e_state["k"] = k + 1
        if k == 0:
            return None
        loss_ready[(k - 1) % 2].synchronize()
        return float(pin_loss[(k - 1) % 2].array[0])

    def e2e_drain():
        k = e2e_state["k"]
        loss_ready[(k - 1) % 2].synchronize()
        return float(pin_loss[(k - 1) % 2].array[0])

    def barrier():
        sk.synchronize()
        if rdv is not None:
            rdv.barrier()

    # nvidia-smi takes ~1 s to come up and holds driver locks while it does: start it before
    # the warm-up so that it is in steady 200 ms polling during the timed regions
    clocks = ClockSampler(env.local_rank)
    if env.rank == 0:
        clocks.start()
        t_wait = time.time()
        while len(clocks.rows) < 3 and time.time() - t_wait < 15.0:   # NVML is up and polling
            time.sleep(0.1)
    # warm-up with the SAME liveness pattern as the timed loop (`last` keeps the previous step's
    # loss -- and through it that step's graph -- alive until the next one has been built, as a
    # user loop `loss = ...` does): the caching allocator reaches its steady state here, not
    # inside the timed region (a cudaMalloc there costs ~100 ms)
    last = None
    for _ in range(args.warmup):
        last = step_resident()
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------
    launches0 = sk.launch_count()
    barrier()
    ev0, ev1 = sk.Event(), sk.Event()
    marks = [sk.Event() for _ in range(args.steps)]
    ev0.record()
    last = None
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        last = step_resident()
        marks[i].record()
    host_issue_ms = (time.perf_counter() - host_t0) * 1e3 / args.steps     # host time to ISSUE a step (no sync inside)
    ev1.record()
    ev1.synchronize()
    barrier()
    ms = ev0.elapsed_ms(ev1)
    per_step = [(ev0 if i == 0 else marks[i - 1]).elapsed_ms(marks[i]) for i in range(args.steps)]
    print(f"[bench rank {env.rank}] per-step ms (timed region): " + " ".join(f"{t:.2f}" for t in per_step)
          + f" | host issue {host_issue_ms:.2f} ms/step", file=sys.stderr, flush=True)
    launches = sk.launch_count() - launches0
    loss_value = last.item()

    # ---- the same K steps again with every kernel launch bracketed by CUDA events on the
    # compute stream: per-family device time and algorithmic work for the roofline object.
    # (Kept apart from the region above so that creating ~2 events per launch does not
    # perturb `value`; profiled_ms_per_step is reported next to ms_per_step.)
    sk.profile_reset()
    sk.profile_enable(True)
    step_resident()            # populates the event pool
    sk.profile_reset()
    barrier()
    pv0, pv1 = sk.Event(), sk.Event()
    pv0.record()
    for _ in range(args.steps):
        step_resident()
    pv1.record()
    pv1.synchronize()
    prof = sk.profile_collect()
    sk.profile_enable(False)
    profiled_ms = pv0.elapsed_ms(pv1)
    barrier()

    # ---- timed region 2: end to end (pinned host -> device every step, loss -> host) -------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = sk.Event(), sk.Event()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e2e_last_loss = e2e_drain()          # the last step's loss reaches the host inside the timed region
    e1.record()
    e1.synchronize()
    barrier()
    ms_e2e = max(e0.elapsed_ms(e1), (time.perf_counter() - t0) * 1e3)   # device events vs host wall clock
    clock_info = clocks.stop() if env.rank == 0 else None

    # ---- data parallel: where the step's time goes (phase events; nsys is not in the image) ----
    dp_timeline = None
    if rdv is not None:
        acc = {"forward": 0.0, "backward": 0.0, "comm_after_backward": 0.0, "optimizer_after_comm": 0.0, "step": 0.0}
        evs = [sk.Event() for _ in range(3)]
        bucket_acc = []          # per bucket: MB, gradients complete / all-reduce done, ms after backward started
        barrier()
        for _ in range(args.steps):
            evs[0].record()
            loss = crit(model(Xd), yd)
            evs[1].record()
            loss.backward()
            evs[2].record()
            ddp.step()
            ev_red, ev_done = ddp.timeline_events()
            ev_done.synchronize()
            acc["forward"] += evs[0].elapsed_ms(evs[1])
            acc["backward"] += evs[1].elapsed_ms(evs[2])
            tail = max(evs[2].elapsed_ms(ev_red), 0.0)         # all-reduce still running after backward ended
            acc["comm_after_backward"] += tail
            acc["optimizer_after_comm"] += max(evs[2].elapsed_ms(ev_done) - tail, 0.0)
            acc["step"] += evs[0].elapsed_ms(ev_done)
            for j, (mb, ready, reduced) in enumerate(ddp.bucket_times(evs[1])):
                if j == len(bucket_acc):
                    bucket_acc.append([mb, 0.0, 0.0])
                bucket_acc[j][1] += ready / args.steps
                bucket_acc[j][2] += reduced / args.steps
        dp_timeline = {k: v / args.steps for k, v in acc.items()}
        dp_timeline["buckets"] = [{"mb": round(mb, 2), "ready_ms": round(a, 3), "reduced_ms": round(b, 3)} for mb, a, b in bucket_acc]
        dp_timeline["note"] = ("ms per step on this rank, CUDA events: forward / backward on the compute stream; "
                               "comm_after_backward = the last bucket's all-reduce finishing after backward's last "
                               "kernel (exposed communication); optimizer_after_comm = the last buckets' Adam + weight "
                               "re-split after that (exposed update); each step drained before the next starts")
        barrier()

    dp_check = None
    if rdv is not None:
        ms = max(rdv.all_gather_float(ms))
        ms_e2e = max(rdv.all_gather_float(ms_e2e))
        # replicated weights must still be replicated: one fp32 sum per parameter (the same
        # reduction kernel on every rank), hashed, compared across ranks
        import hashlib
        ddp.sync_parameters()      # mode p2p with lazy_master: refresh the replicas' fp32 copies first (outside every timed region)
        sums = np.array([float(sk.asnumpy(sk.sum(p._data)).reshape(-1)[0]) for p in params], np.float64)
        digest = hashlib.sha256(sums.tobytes()).hexdigest()[:16]
        digests = rdv.all_gather_str(digest)
        losses = rdv.all_gather_float(float(loss_value))
        dp_check = {"param_checksum": digest, "identical_on_all_ranks": len(set(digests)) == 1,
                    "ranks": len(digests), "local_loss_per_rank": losses}
        if len(set(digests)) != 1:
            print(f"[bench rank {env.rank}] PARAMETER CHECKSUMS DIFFER ACROSS RANKS: {digests}", file=sys.stderr, flush=True)
    ms_per_step = ms / args.steps
    value = batch * env.world / (ms_per_step * 1e-3)
    e2e_value = batch * env.world / (ms_e2e / args.steps * 1e-3)

    if env.rank == 0:
        peaks, peak_kind = load_peaks()
        gemm = prof.get("gemm_tc") or prof.get("gemm_simt") or {"launches": 0, "ms": 0.0, "work": 0.0}
        fam = "gemm_tc" if "gemm_tc" in prof else "gemm_simt"
        achieved = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] else 0.0
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        traffic, traffic_src = load_traffic()
        prep = prof.get("gemm_prep", {"ms": 0.0})
        flop_fams = ("gemm_tc", "gemm_simt")            # work in flops; every other family books bytes
        families = {k: {"launches": v["launches"], "ms_per_step": v["ms"] / args.steps,
                        "rate": (v["work"] / (v["ms"] * 1e-3) / (1e12 if k in flop_fams else 1e9)) if v["ms"] else 0.0,
                        "unit": "TFLOP/s" if k in flop_fams else "GB/s"}
                    for k, v in prof.items()}
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s",
            "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, env.world, batch),
            "e2e": {"value": e2e_value, "unit": "samples/s",
                    "h2d_bytes_per_step": int(Xh.nbytes + yh.nbytes), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "host_issue_ms_per_step": host_issue_ms,
            "clocks": clock_info,
            "roofline": {
                "bound": "tensor", "kernel": fam, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "traffic_source": f"profiles/{traffic_src}: dram__bytes_read.sum + dram__bytes_write.sum of the "
                                  f"8192x4096x4096 forward GEMM launch (algorithmic operand+result bytes: 335.5 MB)"
                                  if traffic_src else None,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_kind}); the path computes fp32-parity "
                               f"GEMMs as fp16x3 on tcgen05 (3 kind::f16 MMAs per K step on fp16 hi/lo operand splits, "
                               f"so at most 1/3 of the 16-bit tensor peak by construction), numerator = 2*M*N*K",
                "mma_rate_tflops": 3.0 * achieved,
                "mma_frac_of_peak": 3.0 * achieved / peak if peak else None,
                "gemm_launches": gemm["launches"], "gemm_ms_per_step": gemm["ms"] / args.steps,
                "share_of_step": gemm["ms"] / profiled_ms if profiled_ms else None,
                "prep_ms_per_step": prep["ms"] / args.steps,
                "measured": "CUDA events around every GEMM kernel launch over the same K steps repeated right after "
                            "the timed region; the fp16 hi/lo operand-split passes are the separate gemm_prep family",
                "profiled_ms_per_step": profiled_ms / args.steps,
            },
            "kernel_families": families,
            "model_flops_per_step": flops_per_step(batch, args.hidden, args.blocks),
            "params": n_params, "final_loss": loss_value, "final_loss_e2e": e2e_last_loss,
        }
        if dp_check is not None:
            line["dp_check"] = dp_check
            line["dp_timeline"] = dp_timeline
        if env.world == 1 and not args.no_cpu_baseline:
            # free the benchmark's device state first: the parity model is a second 272 M-parameter net
            ref = CpuReference(args, install_compat=True)
            ref.set_batch(Xh[:args.cpu_sample_batch], yh[:args.cpu_sample_batch])
            try:
                line["parity_check"] = parity_check(args, ref, sk, soket, nn, Adam)   # = the CPU leg's warm-up step
            except Exception as e:      # the bench line must still be printed
                line["parity_check"] = {"error": f"{type(e).__name__}: {e}"[:300]}
                ref.step()
            sec = ref.time_steps(1, 0)
            line["cpu_baseline"] = {
                "value": args.cpu_sample_batch / sec, "unit": "samples/s", "cores": ref.cores, "kind": ref.kind,
                "blas_threads": blas_threads(),
                "sample": f"1 Adam step of the same model on {args.cpu_sample_batch} rows after 1 warm-up step "
                          f"(full batch is {batch}; the fixed per-step Adam cost is amortised over the sample, full-batch "
                          f"CPU throughput is ~15 % higher); NumPy/OpenBLAS threads = all host cores"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    ddp.close()


# --------------------------------------------------------------------------- drop-in arm
def run_dropin(args, env):
    """The UNMODIFIED reference (oracle/_ref: its Tensor / autodiff / nn / optim code) on
    `soket.gpu()` with soket_b200 in the CuPy seam (compat.install): every one of the ~1500 array
    calls per step is one call into libsoketb200.so, nothing fused above the array layer."""
    if env.rank != 0:
        return
    import soket_b200 as sk
    from oracle import ref_model
    soket = ref_model.import_reference(install_compat=True)
    if soket is None:
        print(json.dumps({"impl": "dropin", "unavailable": "oracle/_ref is not built"}), file=_JSON_OUT, flush=True)
        return
    import soket.nn as rnn
    from soket.nn.init import kaiming_normal
    from soket.optim import Adam
    batch = args.global_batch
    sk.init(env.local_rank)
    X, y = synthetic_batch(batch, 100)
    with soket.gpu():
        np.random.seed(0)
        model = build_model(rnn, args.hidden, args.blocks, norm=args.norm, retain_fn=not args.verbatim_q1)
        for m in model.modules():
            if type(m).__name__ == "Linear":
                kaiming_normal(m.weight)
        opt = Adam(model.parameters(), lr=LR)
        crit = rnn.SoftmaxCrossEntropyLoss()
        model.train(True)
        Xt, yt = soket.Tensor(X), soket.Tensor(y)

        def step():
            loss = crit(model(Xt), yt)
            loss.backward()
            opt.step()
            return loss
        last = None
        for _ in range(args.warmup):
            last = step()
        sk.synchronize()
        n0 = sk.launch_count()
        ev0, ev1 = sk.Event(), sk.Event()
        ev0.record()
        for _ in range(args.steps):
            last = step()
        ev1.record()
        ev1.synchronize()
        ms = ev0.elapsed_ms(ev1) / args.steps
        launches = sk.launch_count() - n0
        loss_value = float(last.item())
    line = {"impl": "dropin", "metric": METRIC, "value": batch / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1, batch), "gpu_launches": int(launches),
            "launches_per_step": launches / args.steps, "final_loss": loss_value,
            "note": "the unmodified reference's own Tensor/autodiff/nn/optim code on soket.gpu(); soket_b200 supplies "
                    "only the array layer (SURVEY.md section 8b seam)"}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "dropin"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the global batch is SPLIT over the ranks (SURVEY.md section 8e, the contract's "
                         "configuration); weak = every rank gets a full batch")
    ap.add_argument("--global-batch", type=int, default=GLOBAL_BATCH, help="rows per step over all GPUs (strong) / per GPU (weak)")
    ap.add_argument("--batch", type=int, default=None, help="alias of --global-batch")
    ap.add_argument("--hidden", type=int, default=HIDDEN)
    ap.add_argument("--blocks", type=int, default=BLOCKS)
    ap.add_argument("--verbatim-q1", action="store_true",
                    help="the model exactly as examples/mlp_resnet/model.py writes it: the blocks' inner layers are "
                         "invisible to parameters() (quirk Q1), 4 trainable tensors / 3.26 M parameters")
    ap.add_argument("--cpu-sample-batch", type=int, default=1024,
                    help="rows per CPU step.  The per-step Adam update (272 M parameters, ~10 s on 8 cores) does not\n"
                         "shrink with the sample, so small samples understate the CPU path: measured on 8 cores\n"
                         "256 rows -> 25, 1024 -> 92, 2048 -> 78, full 8192 -> 109 samples/s (DESIGN.md section 6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--norm", default="layer", choices=["layer", "batch"],
                    help="normalisation of the residual blocks (the bench line is LayerNorm, as examples/mlp_resnet/model.py)")
    args = ap.parse_args()
    if args.batch is not None:
        args.global_batch = args.batch
    _isolate_stdout()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    global NORM
    NORM = args.norm
    env = read_env()
    if args.impl == "reference":
        run_reference(args, env)
    elif args.impl == "dropin":
        run_dropin(args, env)
    else:
        run_ours(args, env)


if __name__ == "__main__":
    main()
