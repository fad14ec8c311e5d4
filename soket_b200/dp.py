"""soket_b200.dp -- single-node data-parallel training (one process per GPU).

New relative to the reference (no collective, no notion of rank: SURVEY.md
section 2 / 8e).  The batch is sharded by contiguous row blocks, parameters and
optimiser state are replicated, and the ONLY collective is an NCCL
all-reduce(sum) of every trainable parameter's gradient, issued on a dedicated
comm stream the moment autodiff finalises that gradient, so it overlaps the rest
of backward; 1/W is folded into the optimiser kernel (`grad_scale`).

Rendezvous: ranks come from the environment torchrun sets (RANK, LOCAL_RANK,
WORLD_SIZE, MASTER_ADDR, MASTER_PORT).  `torch.distributed` is used for exactly
one thing -- a TCPStore to hand the 128-byte NCCL unique id from rank 0 to the
others and for host-side barriers / gathers of a few floats; no tensor ever goes
through it and it is not on the device path.
"""
from __future__ import annotations

import os
from dataclasses import dataclass


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
    """Host-side key/value plumbing over torch.distributed.TCPStore."""

    def __init__(self, env: Env, timeout_s: float = 300.0):
        from datetime import timedelta
        from torch.distributed import TCPStore
        self.env = env
        self._n = 0
        self.store = TCPStore(env.master_addr, env.master_port + 17, env.world, env.rank == 0,
                              timeout=timedelta(seconds=timeout_s), wait_for_workers=True)

    def broadcast_bytes(self, payload: bytes | None, key: str) -> bytes:
        if self.env.rank == 0:
            self.store.set(key, payload)
            return payload
        return bytes(self.store.get(key))

    def barrier(self):
        self._n += 1
        key = f"barrier/{self._n}"
        self.store.add(key, 1)
        import time
        while int(self.store.add(key, 0)) < self.env.world:
            time.sleep(0.0005)

    def all_gather_float(self, value: float) -> list[float]:
        self._n += 1
        self.store.set(f"ag/{self._n}/{self.env.rank}", repr(float(value)))
        return [float(self.store.get(f"ag/{self._n}/{r}").decode()) for r in range(self.env.world)]


def exchange_unique_id(rdv: Rendezvous, make_id) -> bytes:
    """Rank 0 creates the NCCL unique id; every rank returns the same 128 bytes."""
    uid = make_id() if rdv.env.rank == 0 else None
    return rdv.broadcast_bytes(uid, "nccl_uid")


class DataParallel:
    """Gradient all-reduce driver around a soket_b200 model + optimiser.

        dp = DataParallel(optim, rdv)       # after sk.init(local_rank)
        ...
        loss.backward()                     # all-reduces fire per finalised leaf grad
        dp.finish()                         # compute stream waits for the comm stream
        optim.step()                        # grad_scale = 1/W inside the fused kernel
    """

    def __init__(self, optim, rdv: Rendezvous | None, overlap: bool = True):
        from soket_b200 import _fused as F
        from soket_b200 import engine as E
        self.F, self.E = F, E
        self.optim = optim
        self.world = 1 if rdv is None else rdv.env.world
        self.overlap = overlap
        self._ids = {id(p) for p in optim._params}
        if self.world > 1:
            uid = exchange_unique_id(rdv, F.nccl_unique_id)
            F.nccl_init(rdv.env.rank, rdv.env.world, uid)
            optim.grad_scale = 1.0 / self.world
            E.set_leaf_grad_hook(self._on_leaf_grad)

    def broadcast_parameters(self, root: int = 0):
        """Identical initial weights on every rank (section 8d config 5)."""
        if self.world == 1:
            return
        from soket_b200 import _core as B
        for p in self.optim._params:
            if str(p._data.dtype) != "float32":
                continue
            if not p._data.is_contiguous:
                p._data = B.ascontiguousarray(p._data)
            self.F.nccl_broadcast(p._data, root)

    def _on_leaf_grad(self, t):
        if id(t) not in self._ids:
            return
        g = t._grad._data
        if not g.is_contiguous:
            from soket_b200 import _core as B
            g = B.ascontiguousarray(g)
            t._grad._data = g
        self.F.nccl_allreduce(g, self.overlap)

    def finish(self):
        if self.world > 1 and self.overlap:
            self.F.nccl_wait()

    def close(self):
        if self.world > 1:
            self.E.set_leaf_grad_hook(None)
            self.F.nccl_destroy()
