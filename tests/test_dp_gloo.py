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


# ---- the peer-memory mode's protocol (csrc/dp_p2p.cu) with gloo point-to-point transfers ------------------
def _rank_main_sharded(rank, world, port, q):
    """What DataParallel(mode='p2p') does per gradient bucket, with the host logic of soket_b200.dp (arena plan,
    bucket pieces) and gloo send / recv standing in for the copy engines: every rank PULLS its piece of every
    peer's gradient arena, sums the pieces in rank order, runs Adam on its piece only (1/W of the state), and
    PUSHES the piece of the new parameters into every replica."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from oracle import soket_np as O
    from soket_b200 import dp
    env = dp.read_env()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=env.rank, world_size=env.world)
    om = _model()
    names = om.names()
    sizes = [int(om.params[k].size) for k in names]
    offsets, total, buckets = dp.plan_arena(sizes, 600, max_members=32)      # several small buckets
    P = np.zeros(total, np.float32)
    for k, off, n in zip(names, offsets, sizes):
        P[off:off + n] = om.params[k].reshape(-1)
    # Adam on this rank's pieces only: one flat "parameter" per (bucket, piece)
    pieces = []
    for start, end, members in buckets:
        L = dp.shard_len(end - start, world)
        lo = start + rank * L
        pieces.append((start, end, L, lo, max(min(lo + L, end) - lo, 0)))
    opt = O.Adam(len(pieces), lr=1e-2)
    state_elems = sum(p[4] for p in pieces)
    Xs, ys = _data()
    losses = []
    for s in range(STEPS):
        for k, off, n in zip(names, offsets, sizes):
            om.params[k] = P[off:off + n].reshape(om.params[k].shape).copy()
        rows = dp.shard_rows(BATCH, rank, world)
        losses.append(float(om.loss(om.forward(Xs[s][rows]), ys[s][rows])))
        grads = om.backward()
        G = np.zeros(total, np.float32)
        for k, off, n in zip(names, offsets, sizes):
            G[off:off + n] = np.asarray(grads[k], np.float32).reshape(-1)
        newP, oldP, gsum = [], [], []
        for (start, end, L, lo, mylen) in pieces:
            # pull: my piece of every peer's arena (they pull theirs from mine)
            rows_q = {}
            reqs = []
            for q_ in range(world):
                if q_ == rank:
                    continue
                qlo = start + q_ * L
                qlen = max(min(qlo + L, end) - qlo, 0)
                if qlen:
                    reqs.append(dist.isend(torch.from_numpy(G[qlo:qlo + qlen].copy()), q_))
                if mylen:
                    rows_q[q_] = torch.empty(mylen, dtype=torch.float32)
                    reqs.append(dist.irecv(rows_q[q_], q_))
            for r_ in reqs:
                r_.wait()
            acc = None
            for q_ in range(world):                      # rank order whoever owns the piece
                x = G[lo:lo + mylen] if q_ == rank else rows_q[q_].numpy() if mylen else None
                if mylen:
                    acc = x.copy() if acc is None else (acc + x).astype(np.float32)
            gsum.append((acc * np.float32(1.0 / world)).astype(np.float32) if mylen else np.zeros(0, np.float32))
            oldP.append(P[lo:lo + mylen].copy())
        newP = opt.step(oldP, gsum)                      # the reference's Adam arithmetic on 1/W of the elements
        for (start, end, L, lo, mylen), piece in zip(pieces, newP):
            P[lo:lo + mylen] = piece
            reqs = []
            for q_ in range(world):                      # push my piece into every replica, receive theirs
                if q_ == rank:
                    continue
                qlo = start + q_ * L
                qlen = max(min(qlo + L, end) - qlo, 0)
                if mylen:
                    reqs.append(dist.isend(torch.from_numpy(P[lo:lo + mylen].copy()), q_))
                if qlen:
                    buf = torch.empty(qlen, dtype=torch.float32)
                    reqs.append((dist.irecv(buf, q_), qlo, qlen, buf))
            for r_ in reqs:
                if isinstance(r_, tuple):
                    r_[0].wait()
                    P[r_[1]:r_[1] + r_[2]] = r_[3].numpy()
                else:
                    r_.wait()
    dist.barrier()
    q.put((rank, losses, {k: P[off:off + n].reshape(om.params[k].shape).copy() for k, off, n in zip(names, offsets, sizes)},
           state_elems, total))
    dist.destroy_process_group()


def test_two_rank_sharded_adam_protocol_equals_replicated_adam():
    from oracle import soket_np as O
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in out[0][2]:
        assert out[0][2][k].tobytes() == out[1][2][k].tobytes(), k          # replicas stay replicas
    assert out[0][3] + out[1][3] == out[0][4]                    # the ranks' pieces tile the arena exactly ...
    assert max(out[0][3], out[1][3]) <= out[0][4] // 2 + 64 * 40    # ... with ~1/W of the optimizer state per rank
    # the replicated form: both shards' gradients averaged, ONE Adam over whole tensors -- Adam is element-wise,
    # so cutting the arena into pieces changes nothing: bit-identical at W = 2 (one commutative add)
    om = _model()
    names = om.names()
    opt = O.Adam(len(names), lr=1e-2)
    Xs, ys = _data()
    for s in range(STEPS):
        acc = None
        for r in range(2):
            rows = slice(r * BATCH // 2, (r + 1) * BATCH // 2)
            om.loss(om.forward(Xs[s][rows]), ys[s][rows])
            g = om.backward()
            acc = {k: np.asarray(v, np.float32) for k, v in g.items()} if acc is None else \
                {k: (acc[k] + np.asarray(g[k], np.float32)).astype(np.float32) for k in g}
        avg = [(acc[k] * np.float32(0.5)).astype(np.float32) for k in names]
        for k, v in zip(names, opt.step([om.params[k] for k in names], avg)):
            om.params[k] = v
    for k in names:
        assert np.array_equal(out[0][2][k], np.asarray(om.params[k], np.float32)), (k, float(np.abs(out[0][2][k] - om.params[k]).max()))
