"""soket_b200.dp -- single-node data-parallel training (one process per GPU).

New relative to the reference (no collective, no notion of rank: SURVEY.md
section 2 / 8e).  The batch is sharded by contiguous row blocks, parameters and
optimiser state are replicated, and the ONLY collective is an NCCL
all-reduce(sum) of the gradients; 1/W is folded into the optimiser kernel
(`grad_scale`).

How a step runs at W > 1:

  * all gradients live in ONE flat fp32 arena, laid out in REVERSE parameter order
    (the order backward finalises them).  Each parameter's slot is handed to the
    fused backward kernels (`Tensor._grad_buf`): the dW GEMM, the bias-gradient
    reduction and the LayerNorm parameter reductions write straight into it, so
    there is no gather copy (a gradient produced elsewhere is copied in).
  * the arena is cut into contiguous BUCKETS (~64 MiB, i.e. about one Linear layer
    each).  The moment autodiff has finalised the last gradient of a bucket, one
    `ncclAllReduce` over that bucket's range is queued on the comm stream (after an
    event on the compute stream), and the optimizer update of exactly those
    parameters is queued on the optimizer stream behind the all-reduce's event --
    both overlap the rest of backward, which keeps the compute stream.
  * `step()` flushes whatever is still pending, advances the optimizer's per-step
    state and makes the compute stream wait for the optimizer stream.

Updating a layer's parameters while backward is still running is safe: every kernel
that reads them in this step (the layer's own forward and backward) was queued on the
compute stream before the event its bucket waits for.

Rendezvous: ranks come from the environment torchrun sets (RANK, LOCAL_RANK,
WORLD_SIZE, MASTER_ADDR, MASTER_PORT).  The 128-byte NCCL unique id goes from rank 0 to the
others through a file under /dev/shm (class Rendezvous), as do host-side barriers and the few
floats the benchmark gathers; no tensor ever goes through it, it is not on the device path, and
neither torch nor torch.distributed is imported by a training process.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field


@dataclass
class Env:
    rank: int
    local_rank: int
    world: int
    master_addr: str
    master_port: int


def read_env() -> Env:
    return Env(
        rank=int(os.environ.get("RANK", "0")),
        local_rank=int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))),
        world=int(os.environ.get("WORLD_SIZE", "1")),
        master_addr=os.environ.get("MASTER_ADDR", "127.0.0.1"),
        master_port=int(os.environ.get("MASTER_PORT", "29500")),
    )


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous, equal row shards: rank r owns [r*n/W, (r+1)*n/W) (section 8e)."""
    if n_rows % world != 0:
        raise ValueError(f"batch of {n_rows} rows does not split evenly over {world} ranks")
    per = n_rows // world
    return slice(rank * per, (rank + 1) * per)


class Rendezvous:
    """Host-side key/value plumbing between the ranks of ONE node: small files under /dev/shm.

    Nothing but the 128-byte NCCL id, barriers and a few floats for reporting ever goes through it
    ("a /dev/shm file"); no tensor does, it is not on the device path, and the
    process imports neither torch nor a socket library for it.  The directory is keyed by the
    rendezvous port and the launcher's pid (all ranks of a torchrun job share their parent), rank 0
    creates it afresh and removes it on `close()` / interpreter exit.  Values are written to a
    temporary name and renamed, so a reader sees a whole value or none."""

    def __init__(self, env: Env, timeout_s: float = 300.0):
        import atexit
        import shutil
        import tempfile
        self.env = env
        self.timeout_s = float(timeout_s)
        self._n = 0
        base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        job = os.environ.get("SOKET_B200_RDV_ID") or f"{env.master_port}_{os.getppid()}"
        self.dir = os.path.join(base, f"soket_b200_rdv_{job}")
        self._closed = False
        if env.rank == 0:
            shutil.rmtree(self.dir, ignore_errors=True)       # a crashed earlier job with the same key
            os.makedirs(self.dir + ".new", exist_ok=True)
            os.rename(self.dir + ".new", self.dir)
        else:
            self._wait(lambda: os.path.isdir(self.dir), "the rendezvous directory of rank 0")
        atexit.register(self.close)

    def _wait(self, ready, what):
        import time
        t0 = time.monotonic()
        pause = 0.0002
        while not ready():
            if time.monotonic() - t0 > self.timeout_s:
                raise TimeoutError(f"soket_b200.dp rendezvous: rank {self.env.rank} waited {self.timeout_s:.0f} s for {what} "
                                   f"in {self.dir}")
            time.sleep(pause)
            pause = min(pause * 1.5, 0.01)

    def _put(self, key: str, payload: bytes):
        path = os.path.join(self.dir, key)
        tmp = f"{path}.tmp{self.env.rank}"
        with open(tmp, "wb") as f:
            f.write(payload)
        os.rename(tmp, path)

    def _get(self, key: str) -> bytes:
        path = os.path.join(self.dir, key)
        self._wait(lambda: os.path.exists(path), f"key {key!r}")
        with open(path, "rb") as f:
            return f.read()

    def broadcast_bytes(self, payload: bytes | None, key: str) -> bytes:
        """Rank 0's `payload` on every rank.  Collective: the n-th call on every rank is one exchange."""
        self._n += 1
        name = f"bc_{self._n}_{key}"
        if self.env.rank == 0:
            self._put(name, payload)
            return payload
        return self._get(name)

    def all_gather_bytes(self, payload: bytes) -> list[bytes]:
        self._n += 1
        self._put(f"ag_{self._n}_{self.env.rank}", payload)
        return [self._get(f"ag_{self._n}_{r}") for r in range(self.env.world)]

    def all_gather_str(self, value: str) -> list[str]:
        return [b.decode() for b in self.all_gather_bytes(value.encode())]

    def barrier(self):
        self.all_gather_str("")

    def all_gather_float(self, value: float) -> list[float]:
        return [float(v) for v in self.all_gather_str(repr(float(value)))]

    def close(self):
        """Rank 0 removes the directory (call after a final barrier; also runs at interpreter exit)."""
        if self._closed:
            return
        self._closed = True
        try:
            self._put(f"done_{self.env.rank}", b"")
        except OSError:
            return
        if self.env.rank == 0:
            import shutil
            import time
            t0 = time.monotonic()       # the others may still be reading the last exchange
            while (time.monotonic() - t0 < 10.0
                   and not all(os.path.exists(os.path.join(self.dir, f"done_{r}")) for r in range(self.env.world))):
                time.sleep(0.002)
            shutil.rmtree(self.dir, ignore_errors=True)


def exchange_unique_id(rdv: Rendezvous, make_id) -> bytes:
    """Rank 0 creates the NCCL unique id; every rank returns the same 128 bytes."""
    uid = make_id() if rdv.env.rank == 0 else None
    return rdv.broadcast_bytes(uid, "nccl_uid")


SLOT_ALIGN = 64           # floats: every slot starts on a 256-byte boundary (vector loads, TMA-free GEMM epilogue)


def plan_arena(sizes, bucket_floats):
    """Layout of the flat gradient arena for parameters of `sizes` elements (in parameter-list
    order).  Slots are placed in REVERSE order -- backward reaches the last layer first -- each
    aligned to SLOT_ALIGN floats, and cut greedily into buckets of at least `bucket_floats` floats.
    Returns (offsets, total, buckets) with offsets[i] the slot of parameter i and
    buckets = [(start, end, [parameter indices in slot order])].  Pure host logic (CPU-tested)."""
    offsets = [0] * len(sizes)
    buckets = []
    off, start, members = 0, 0, []
    for i in reversed(range(len(sizes))):
        offsets[i] = off
        off += (int(sizes[i]) + SLOT_ALIGN - 1) // SLOT_ALIGN * SLOT_ALIGN
        members.append(i)
        if off - start >= bucket_floats:
            buckets.append((start, off, members))
            start, members = off, []
    if members:
        buckets.append((start, off, members))
    return offsets, off, buckets


@dataclass
class _Bucket:
    start: int
    end: int
    members: list
    view: object = None
    ev_ready: object = None
    ev_reduced: object = None
    arrived: int = 0
    launched: bool = False
    seen: set = field(default_factory=set)


class DataParallel:
    """Bucketed gradient all-reduce + per-bucket optimizer update around a soket_b200 optimiser.

        ddp = DataParallel(optim, rdv)      # after sk.init(local_rank)
        ...
        loss.backward()                     # buckets are reduced and applied as they complete
        ddp.step()                          # flush, advance the optimizer, join the streams

    At world size 1 `step()` is `optim.step()`.  While a DataParallel object is open the gradients
    of the optimizer's parameters live in its arena: `p.grad` of one step is overwritten by the
    next backward, and after `step()` it holds the SUM over ranks (the 1/W is applied inside the
    optimizer kernel)."""

    def __init__(self, optim, rdv: Rendezvous | None, overlap: bool = True, bucket_mb: float | None = None):
        from soket_b200 import _core as B
        from soket_b200 import _fused as F
        from soket_b200 import engine as E
        self.B, self.F, self.E = B, F, E
        self.optim = optim
        self.world = 1 if rdv is None else rdv.env.world
        self.overlap = overlap
        self._closed = False
        if self.world == 1:
            return
        if bucket_mb is None:
            bucket_mb = float(os.environ.get("SOKET_B200_DP_BUCKET_MB", "64"))
        params = optim._params
        for p in params:
            if str(p._data.dtype) != "float32":
                raise TypeError("data-parallel training supports float32 parameters only "
                                f"(got {p._data.dtype} of shape {p.shape})")
            if not p._data.is_contiguous:
                p._data = B.ascontiguousarray(p._data)
        uid = exchange_unique_id(rdv, F.nccl_unique_id)
        F.nccl_init(rdv.env.rank, rdv.env.world, uid)
        optim.grad_scale = 1.0 / self.world
        offsets, total, buckets = plan_arena([int(p.size) for p in params], int(bucket_mb * (1 << 20) / 4))
        self.arena = B.zeros((max(total, 1),), "float32")      # padding between slots stays zero
        self._slots = []
        for p, off in zip(params, offsets):
            slot = self.arena[off:off + int(p.size)].reshape(p.shape)
            self._slots.append(slot)
            p._grad_buf = slot
        self._buckets = []
        self._where = {}
        for start, end, members in buckets:
            b = _Bucket(start, end, members, self.arena[start:end], B.Event(), B.Event())
            for i in members:
                self._where[id(params[i])] = (len(self._buckets), i)
            self._buckets.append(b)
        self._ev_done = B.Event()
        self._last_bucket = None
        E.set_leaf_grad_hook(self._on_leaf_grad)

    # ------------------------------------------------------------------------------------------
    def broadcast_parameters(self, root: int = 0):
        """Identical initial weights on every rank (section 8d config 5)."""
        if self.world == 1:
            return
        for p in self.optim._params:
            self.F.nccl_broadcast(p._data, root)

    def _on_leaf_grad(self, t):
        ent = self._where.get(id(t))
        if ent is None:
            return
        bi, i = ent
        slot = self._slots[i]
        g = t._grad._data
        if g.data_ptr != slot.data_ptr:
            # produced outside the fused backward kernels (or summed from several partials)
            slot[...] = g if g.shape == slot.shape else g.reshape(slot.shape)
            t._grad._data = slot
        b = self._buckets[bi]
        if i not in b.seen:
            b.seen.add(i)
            b.arrived += 1
        if self.overlap and not b.launched and b.arrived == len(b.members):
            self._launch(b)

    def _launch(self, b):
        """Queue the bucket's all-reduce (comm stream) and its optimizer update (optimizer stream)."""
        B = self.B
        b.ev_ready.record(B.STREAM_COMPUTE)
        b.ev_ready.wait(B.STREAM_COMM)               # the gradients of this bucket are complete
        self.F.nccl_allreduce_on(b.view, B.STREAM_COMM)
        b.ev_reduced.record(B.STREAM_COMM)
        b.ev_reduced.wait(B.STREAM_OPT)
        B.launch_stream(B.STREAM_OPT)
        try:
            idx = sorted(b.seen)
            self.optim.update(idx)
            # the updated weights' GEMM operand splits, off the critical path too: the next forward
            # finds them cached (engine.linear -> get_split)
            if self.E.presplit_enabled():
                for i in idx:
                    w = self.optim._params[i]._data
                    # only with the |max| the optimizer kernel left behind: the split then needs no
                    # scratch allocation, which the stream-ordered allocator could hand to the
                    # compute stream while this stream still uses it
                    if self.E.weight_split_eligible(w) and B.has_absmax(w):
                        B.get_split(w)
        finally:
            B.launch_stream(B.STREAM_COMPUTE)
        b.launched = True
        self._last_bucket = b

    def step(self):
        """The optimizer step of a data-parallel iteration (call after `backward()`)."""
        if self.world == 1:
            self.optim.step()
            return
        B = self.B
        for b in self._buckets:
            if not b.launched and b.arrived:
                self._launch(b)
        B.launch_stream(B.STREAM_OPT)
        try:
            self.optim.end_step()
        finally:
            B.launch_stream(B.STREAM_COMPUTE)
        self._ev_done.record(B.STREAM_OPT)
        self._ev_done.wait(B.STREAM_COMPUTE)         # the next forward reads the updated parameters
        for b in self._buckets:
            b.arrived, b.launched = 0, False
            b.seen.clear()

    def timeline_events(self):
        """(event after the last all-reduce of the step, event after the last optimizer update) --
        recorded on the comm / optimizer streams by the most recent `step()`; for phase timing."""
        return (self._last_bucket.ev_reduced if self._last_bucket is not None else None), self._ev_done

    def finish(self):
        """Kept for callers of the round-1 API (`finish(); optim.step()`): reduce whatever is
        pending WITHOUT applying it; the caller's `optim.step()` then updates on the compute stream."""
        if self.world == 1:
            return
        raise RuntimeError("DataParallel.finish() was replaced by DataParallel.step(), which also runs the "
                           "optimizer (per bucket, overlapped with backward)")

    def close(self):
        if self.world > 1 and not self._closed:
            self._closed = True
            self.E.set_leaf_grad_hook(None)
            for p in self.optim._params:
                p._grad_buf = None
            self.B.synchronize()
            self.F.nccl_destroy()
