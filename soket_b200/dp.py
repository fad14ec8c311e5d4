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

Mode 'p2p' (opt-in: `DataParallel(mode="p2p")` / SOKET_B200_DP_MODE=p2p; Adam) replaces the
all-reduce + replicated update of a bucket by reduce-scatter + sharded Adam + all-gather over
CUDA-IPC peer memory: the copy engines pull this rank's piece of every peer's gradients, one
local kernel sums them in rank order, applies Adam to 1/W of the state and emits the GEMM
weights' fp16 hi / lo operand split, the copy engines push the piece into every replica's
arenas; ready / done flags in peer memory order it (csrc/dp_p2p.cu).
Measured: +5.5 % over the NCCL mode at 2 ranks, slower at 8 (profiles/r2_dp_scaling.md).

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


def shard_len(bucket_len: int, world: int) -> int:
    """Elements per rank of a bucket under the peer-memory update: ceil(len / world) rounded up to 64 elements
    (mirrors sk_p2p_shard_len)."""
    if bucket_len <= 0:
        return 0
    per = (bucket_len + world - 1) // world
    return (per + 63) // 64 * 64


def piece_of(bucket_start: int, bucket_len: int, offset: int, n: int, rank: int, world: int):
    """(start, count): the part of a tensor of n elements at arena offset `offset` that falls into rank's
    contiguous piece of the bucket [bucket_start, bucket_start + bucket_len).  Pure host logic (CPU-tested)."""
    L = shard_len(bucket_len, world)
    lo = bucket_start + rank * L
    hi = min(lo + L, bucket_start + bucket_len)
    a, b = max(lo, offset), min(hi, offset + n)
    return (a - offset, b - a) if b > a else (0, 0)


def exchange_unique_id(rdv: Rendezvous, make_id) -> bytes:
    """Rank 0 creates the NCCL unique id; every rank returns the same 128 bytes."""
    uid = make_id() if rdv.env.rank == 0 else None
    return rdv.broadcast_bytes(uid, "nccl_uid")


SLOT_ALIGN = 64           # floats: every slot starts on a 256-byte boundary (vector loads, TMA-free GEMM epilogue)


def plan_arena(sizes, bucket_floats, max_members=None):
    """Layout of the flat gradient arena for parameters of `sizes` elements (in parameter-list
    order).  Slots are placed in REVERSE order -- backward reaches the last layer first -- each
    aligned to SLOT_ALIGN floats, and cut greedily into buckets of at least `bucket_floats` floats.
    Returns (offsets, total, buckets) with offsets[i] the slot of parameter i and
    buckets = [(start, end, [parameter indices in slot order])]; a bucket also closes at `max_members`
    tensors (the peer-memory update kernel takes 32 per launch).  Pure host logic (CPU-tested)."""
    offsets = [0] * len(sizes)
    buckets = []
    off, start, members = 0, 0, []
    for i in reversed(range(len(sizes))):
        offsets[i] = off
        off += (int(sizes[i]) + SLOT_ALIGN - 1) // SLOT_ALIGN * SLOT_ALIGN
        members.append(i)
        if off - start >= bucket_floats or (max_members is not None and len(members) >= max_members):
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
    t_launch: int = -1
    index: int = -1
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

    def __init__(self, optim, rdv: Rendezvous | None, overlap: bool = True, bucket_mb: float | None = None,
                 mode: str | None = None, share_grads: bool | None = None, lazy_master: bool | None = None):
        from soket_b200 import _core as B
        from soket_b200 import _fused as F
        from soket_b200 import engine as E
        self.B, self.F, self.E = B, F, E
        self.optim = optim
        self.world = 1 if rdv is None else rdv.env.world
        self.rdv = rdv
        self.mode = "nccl"
        self.overlap = overlap
        # SOKET_B200_DP_OPT_OVERLAP=0: one optimizer launch after the last all-reduce instead of one per bucket
        # beside backward (the HBM-bound update then does not compete with the GEMMs for bandwidth)
        self.opt_overlap = os.environ.get("SOKET_B200_DP_OPT_OVERLAP", "1") != "0"
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
        mode = mode or os.environ.get("SOKET_B200_DP_MODE", "nccl")
        if mode not in ("nccl", "p2p"):
            raise ValueError(f"DataParallel: unknown mode {mode!r} (nccl | p2p)")
        if mode == "p2p":
            why = self._p2p_unsupported()
            if why:
                raise ValueError(f"DataParallel(mode='p2p'): {why}")
        self.mode = mode
        if share_grads is None:
            share_grads = os.environ.get("SOKET_B200_DP_SHARE_GRADS", "0") == "1"
        self.share_grads = bool(share_grads)
        # mode 'p2p': GEMM weights reach the other replicas as their fp16 hi / lo operand split only (what forward
        # and backward read); a replica's fp32 copy of a weight is current on its owner's piece only until
        # `sync_parameters()` (called by `close()`) -- halves the bytes pushed over NVLink per step
        if lazy_master is None:
            lazy_master = os.environ.get("SOKET_B200_DP_LAZY_MASTER", "0") == "1"
        self.lazy_master = bool(lazy_master) and mode == "p2p"
        offsets, total, buckets = plan_arena([int(p.size) for p in params], int(bucket_mb * (1 << 20) / 4),
                                             max_members=32 if mode == "p2p" else None)
        self._offsets, self._total = offsets, total
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
            b.index = len(self._buckets)
            for i in members:
                self._where[id(params[i])] = (len(self._buckets), i)
            self._buckets.append(b)
        self._ev_done = B.Event()
        self._last_bucket = None
        self._n_launch = 0
        if mode == "p2p":
            self._p2p_setup()
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

    # ------------------------------------------------------------------------------------------
    # mode 'p2p': reduce-scatter + Adam + operand split + all-gather as one kernel per bucket over NVLink peer
    # memory (csrc/dp_p2p.cu).  The parameters move into ONE arena per rank (same layout as the gradient
    # arena) that the peers map through CUDA IPC, the GEMM weights' fp16 hi / lo splits into two more; every
    # rank keeps the Adam moments of its shard of every tensor only.
    def _p2p_unsupported(self):
        from soket_b200.optim import Adam, adam_ratio_bound
        o = self.optim
        if not isinstance(o, Adam):
            return "the fused peer-memory update implements Adam (use mode='nccl' for other optimizers)"
        if o._capturable:
            return "Adam(capturable=True) is not supported"
        if self.world > 8:
            return "at most 8 ranks (one NVSwitch domain)"
        if adam_ratio_bound(o._beta1, o._beta2, 1) is None:
            return "no finite bound of |m_hat / sqrt(v_hat)| for these betas"
        if any(u is not None for u in o._u):
            return "the optimizer has already taken steps (its state is replicated)"
        return None

    def _p2p_setup(self):
        import pickle
        import numpy as np
        B, F, E = self.B, self.F, self.E
        params, rdv = self.optim._params, self.rdv
        rank, world = rdv.env.rank, self.world
        if len(self._buckets) > 256:
            raise ValueError(f"DataParallel(mode='p2p'): {len(self._buckets)} buckets (at most 256): raise bucket_mb")
        total = max(self._total, 1)
        self._parena = B.zeros((total,), "float32")
        self._hi = B.zeros((total,), "float16")
        self._lo = B.zeros((total,), "float16")
        nb, ns = len(self._buckets), len(params)
        self._flag_words = 2 * nb * 8 + 2 * ns * 8
        self._flags = B.zeros((self._flag_words,), "uint32")
        self._split = [None] * ns
        self._m, self._v, self._shard, self._fresh = [None] * ns, [None] * ns, [None] * ns, [True] * ns
        for i, (p, off) in enumerate(zip(params, self._offsets)):
            view = B.arena_view(self._parena, off, tuple(p.shape))
            view[...] = p._data
            p._data = view                                      # the tensor's storage IS its window of the arena
        stage = 0
        for b in self._buckets:
            stage = max(stage, world * shard_len(b.end - b.start, world))
            for i in b.members:
                start, count = piece_of(b.start, b.end - b.start, self._offsets[i], int(params[i].size), rank, world)
                self._shard[i] = (start, count)
                if count:
                    self._m[i] = B.zeros((count,), "float32")
                    self._v[i] = B.zeros((count,), "float32")
        self._staging = B.zeros((max(stage, 1),), "float32")    # world rows of one bucket shard (buckets run in order)
        B.synchronize()
        mine = [F.ipc_export(a) for a in (self.arena, self._parena, self._hi, self._lo, self._flags)]
        everyone = [pickle.loads(x) for x in rdv.all_gather_bytes(pickle.dumps(mine))]
        own = [a.data_ptr for a in (self.arena, self._parena, self._hi, self._lo, self._flags)]
        addr = [[own[k] if q == rank else F.ipc_open(*everyone[q][k]) for q in range(world)] for k in range(5)]
        self._peers = F.P2pPeers(world, rank, nb, ns, addr[0], addr[1], addr[2], addr[3], addr[4], self._flags)
        self._p2p_ready = False
        self._step = 1
        rdv.barrier()                                           # every rank has mapped every arena

    def _p2p_prepare(self):
        """Before the first update (and after `broadcast_parameters`): every GEMM weight's operand split moves
        into the hi / lo arenas and its |max| seeds the table the kernels derive the split scale from."""
        import numpy as np
        B, E = self.B, self.E
        params = self.optim._params
        parts = np.zeros((len(params), 8), np.uint32)
        if E.presplit_enabled():
            for i, (p, off) in enumerate(zip(params, self._offsets)):
                w = p._data
                if not E.weight_split_eligible(w):
                    continue
                n = int(p.size)
                hi = B.arena_view(self._hi, off, tuple(p.shape))
                lo = B.arena_view(self._lo, off, tuple(p.shape))
                sm = B.split_f16(w, out_hi=hi, out_lo=lo)       # binds itself to w
                self._split[i] = sm
                amax = np.float32(B.asnumpy(sm.scale)[2])
                parts[i, :] = amax.view(np.uint32)
        nb = len(self._buckets)
        base = 2 * nb * 8 + len(params) * 8                      # parts[1]: read by step 1
        self._flags[base:base + parts.size] = B.array(parts.reshape(-1))
        self._p2p_ready = True

    def _launch_p2p(self, b):
        B, F = self.B, self.F
        if not self._p2p_ready:
            self._p2p_prepare()
        o = self.optim
        from soket_b200.optim import adam_ratio_bound
        params = o._params
        if len(b.seen) != len(b.members):
            raise RuntimeError("DataParallel(mode='p2p'): a parameter of this bucket received no gradient in this step "
                               "(every rank must update the same tensors); use mode='nccl' for such models")
        tensors = []
        for i in b.members:
            p = params[i]
            if p._data.data_ptr != self._peers_param_ptr(i):
                raise RuntimeError("DataParallel(mode='p2p'): a parameter's storage was rebound outside the arena "
                                   "(assign through p.data[...] = value, or close() the DataParallel object first)")
            start, count = self._shard[i]
            tensors.append((p._data, self._offsets[i], start, count, self._m[i], self._v[i], self._split[i], i, self._fresh[i]))
            self._fresh[i] = False
        rb = adam_ratio_bound(o._beta1, o._beta2, o._t)
        ub = abs(float(o._lr)) * rb * 1.0001
        b.ev_ready.record(B.STREAM_COMPUTE)
        b.ev_ready.wait(B.STREAM_OPT)                # the gradients of this bucket are complete on this rank
        B.launch_stream(B.STREAM_OPT)
        try:
            F.dp_p2p_update(self._peers, b.index, self._step, b.start, b.end - b.start, self._staging, tensors, o._lr, o._beta1, o._beta2, o._eps,
                            o._weight_decay if o._have_weight_decay else 0.0, o._one_minus_beta1_t, o._one_minus_beta2_t,
                            1.0 / self.world, ub, self.share_grads, self.lazy_master)
        finally:
            B.launch_stream(B.STREAM_COMPUTE)
        b.ev_reduced.record(B.STREAM_OPT)            # this rank's shard of the bucket is updated and published
        b.launched = True
        self._n_launch += 1
        b.t_launch = self._n_launch
        self._last_bucket = b

    def _peers_param_ptr(self, i):
        return self._parena.data_ptr + 4 * self._offsets[i]

    def _launch(self, b):
        """Queue the bucket's all-reduce (comm stream) and its optimizer update (optimizer stream)."""
        if self.mode == "p2p":
            return self._launch_p2p(b)
        B = self.B
        b.ev_ready.record(B.STREAM_COMPUTE)
        b.ev_ready.wait(B.STREAM_COMM)               # the gradients of this bucket are complete
        self.F.nccl_allreduce_on(b.view, B.STREAM_COMM)
        b.ev_reduced.record(B.STREAM_COMM)
        b.launched = True
        self._n_launch += 1
        b.t_launch = self._n_launch
        self._last_bucket = b
        if self.opt_overlap:
            self._update(b, sorted(b.seen))

    def _update(self, b, idx):
        """The optimizer update of parameters `idx` on the optimizer stream, after bucket b's all-reduce."""
        B = self.B
        b.ev_reduced.wait(B.STREAM_OPT)
        B.launch_stream(B.STREAM_OPT)
        try:
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

    def step(self):
        """The optimizer step of a data-parallel iteration (call after `backward()`)."""
        if self.world == 1:
            self.optim.step()
            return
        B = self.B
        for b in self._buckets:
            if not b.launched and b.arrived:
                self._launch(b)
        if self.mode == "p2p":
            self.optim.end_step()
            # the next forward reads weights every OWNER rank has written: wait for all of them (their `done`
            # also says they have finished reading this rank's gradient slots)
            self.F.dp_p2p_wait(self._peers, self._step, [b.index for b in self._buckets if b.launched])
            self._step += 1
            self._ev_done.record(B.STREAM_COMPUTE)
            for b in self._buckets:
                b.arrived, b.launched = 0, False
                b.seen.clear()
            return
        if not self.opt_overlap and self._last_bucket is not None:
            # NCCL collectives of one communicator complete in issue order: the last one covers them all
            self._update(self._last_bucket, sorted(i for b in self._buckets if b.launched for i in b.seen))
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

    def sync_parameters(self):
        """Collective (every rank calls it, between steps): after it every replica's fp32 parameters are current.
        A no-op unless mode 'p2p' runs with lazy_master."""
        if self.world == 1 or not self.lazy_master or self._closed:
            return
        self.B.synchronize()
        self.rdv.barrier()                           # nobody is still inside a step
        for b in self._buckets:
            self.F.dp_p2p_gather(self._peers, b.start, b.end - b.start)
        self.B.synchronize()
        self.rdv.barrier()

    def bucket_times(self, origin):
        """[(MB, ms from `origin` until the bucket's gradients were complete, ms until its all-reduce had
        finished)] of the most recent step, in launch order (call after the step has drained)."""
        out = []
        for b in sorted((b for b in self._buckets if b.t_launch >= 0), key=lambda b: b.t_launch):
            out.append(((b.end - b.start) * 4 / 1e6, origin.elapsed_ms(b.ev_ready), origin.elapsed_ms(b.ev_reduced)))
        return out

    def finish(self):
        """Kept for callers of the round-1 API (`finish(); optim.step()`): reduce whatever is
        pending WITHOUT applying it; the caller's `optim.step()` then updates on the compute stream."""
        if self.world == 1:
            return
        raise RuntimeError("DataParallel.finish() was replaced by DataParallel.step(), which also runs the "
                           "optimizer (per bucket, overlapped with backward)")

    def close(self):
        if self.world > 1 and not self._closed:
            self.sync_parameters()
            self._closed = True
            self.E.set_leaf_grad_hook(None)
            for p in self.optim._params:
                p._grad_buf = None
            self.B.synchronize()
            if self.mode == "p2p":
                self.rdv.barrier()                   # nobody unmaps while a peer's kernel may still write here
                self.F.ipc_close_all()
                self.rdv.barrier()
            self.F.nccl_destroy()
