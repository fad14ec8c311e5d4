"""CPU, world_size 2, gloo: the data-parallel CONTRACT of soket_b200.dp
with a real collective between two processes -- contiguous row shards per rank
(dp.shard_rows / utils.data.shard_bounds), local gradients of the local-mean loss,
all-reduce(sum) of every gradient, 1/W folded into the optimizer step (`grad_scale`) --
reproduces the single-process step on the global batch.  The arithmetic is the oracle's (NumPy);
on the GPU the same protocol runs over NCCL (tests/test_dp_gpu.py)."""
import multiprocessing as mp
import os
import socket

import numpy as np

DIM, HID, NB, C, BATCH, STEPS = 32, 16, 2, 10, 48, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    from oracle import soket_np as O
    om = O.MLPResNet(DIM, HID, NB, C, norm="layer")
    rng = np.random.default_rng(0)
    for k in om.names():
        shp = om.params[k].shape
        if k.endswith(".W"):
            om.params[k] = (rng.standard_normal(shp) * np.sqrt(2.0 / shp[0])).astype("float32")
        elif k.endswith(".g"):
            om.params[k] = (1 + 0.1 * rng.standard_normal(shp)).astype("float32")
        else:
            om.params[k] = (0.1 * rng.standard_normal(shp)).astype("float32")
    return om


def _data():
    rng = np.random.default_rng(1)
    return (rng.random((STEPS, BATCH, DIM), dtype=np.float32), rng.integers(0, C, (STEPS, BATCH)).astype(np.uint8))


def _rank_main(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from oracle import soket_np as O
    from soket_b200 import dp
    from soket_b200.utils.data import shard_bounds
    env = dp.read_env()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=env.rank, world_size=env.world)
    om = _model()
    opt = O.SGD(len(om.names()), lr=0.05)
    Xs, ys = _data()
    losses = []
    for s in range(STEPS):
        rows = dp.shard_rows(BATCH, env.rank, env.world)
        assert (rows.start, rows.stop) == shard_bounds(BATCH, env.rank, env.world)
        logits = om.forward(Xs[s][rows])
        losses.append(float(om.loss(logits, ys[s][rows])))      # mean over the LOCAL shard
        grads = om.backward()
        names = om.names()
        for k in names:                                          # the only collective: sum of gradients
            t = torch.from_numpy(np.ascontiguousarray(grads[k]))
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            grads[k] = (t.numpy() * np.float32(1.0 / env.world)).astype("float32")   # grad_scale = 1/W
        new = opt.step([om.params[k] for k in names], [grads[k] for k in names])
        for k, v in zip(names, new):
            om.params[k] = v
    dist.barrier()
    q.put((rank, losses, {k: om.params[k] for k in om.names()}))
    dist.destroy_process_group()


def test_two_rank_gloo_training_equals_global_batch_training():
    from oracle import soket_np as O
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas stay identical: same parameters on both ranks, bit for bit
    for k in out[0][2]:
        assert out[0][2][k].tobytes() == out[1][2][k].tobytes(), k
    # ... and equal to one process training on the global batch (LayerNorm: per-sample
    # statistics, so the shard-mean gradients average to the global-mean gradient)
    om = _model()
    opt = O.SGD(len(om.names()), lr=0.05)
    Xs, ys = _data()
    for s in range(STEPS):
        want, _ = om.train_step(Xs[s], ys[s], opt)
        got = 0.5 * (out[0][1][s] + out[1][1][s])               # the bench averages the shard losses
        assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), (s, got, want)
    for k in om.names():
        ref = om.params[k]
        err = np.abs(out[0][2][k] - ref).max() / max(np.abs(ref).max(), 1e-30)
        assert err <= 1e-5, (k, err)
