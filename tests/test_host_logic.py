"""CPU tests of host-side logic: dtype promotion table, scalar typing, DP sharding /
rendezvous (world_size 2 over a TCPStore, gloo-free), bench helpers."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest


def test_promotion_table_matches_reference(ref_soket):
    import soket_b200.engine as E
    soket = ref_soket
    names = E.DType._names
    for a in names:
        for b in names:
            want = soket.promote_types(getattr(soket, a), getattr(soket, b)).name
            got = E.promote_types(E._DTYPES[a], E._DTYPES[b]).name
            assert got == want, (a, b, got, want)


def test_scalar_dtypes_and_dtype_api():
    import soket_b200.engine as E
    assert E._scalar_dtype(1) is E.int32 and E._scalar_dtype(1.0) is E.float32 and E._scalar_dtype(True) is E.bool_
    with pytest.raises(ValueError):
        E.DType("complex64")
    assert E.DType("float32") == E.float32 and str(E.float32) == "float32"
    assert E.promote_types(E.bool_, E.uint8) is E.uint8
    assert E.promote_types(E.int64, E.uint64).name == "float32"   # reference table, not NumPy's float64


def test_shard_rows_partitions_the_batch():
    from soket_b200.dp import shard_rows
    for n, w in ((8192, 1), (8192, 2), (8192, 8), (100, 4)):
        seen = np.zeros(n, int)
        for r in range(w):
            seen[shard_rows(n, r, w)] += 1
        assert np.all(seen == 1)
    with pytest.raises(ValueError):
        shard_rows(100, 0, 3)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from soket_b200 import dp
    env = dp.read_env()
    rdv = dp.Rendezvous(env, timeout_s=60)
    uid = dp.exchange_unique_id(rdv, lambda: bytes(range(128)))
    rdv.barrier()
    vals = rdv.all_gather_float(1.5 + rank)
    rdv.barrier()
    q.put((rank, uid, vals, dp.shard_rows(64, env.rank, env.world)))


def test_rendezvous_world_size_2():
    """The N>1 host path (unique-id hand-off, barrier, scalar gather) on CPU."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == out[1][1] == bytes(range(128))
    assert out[0][2] == out[1][2] == [1.5, 2.5]
    assert out[0][3] == slice(0, 32) and out[1][3] == slice(32, 64)


def test_public_api_covers_the_reference(ref_soket):
    """Every public name a Soket script can reach -- soket.*, soket.nn.*, soket.optim.*,
    soket.nn.init / functional, Tensor / Module / optimizer attributes -- exists here
    (submodule names and imports the reference leaks through `import *` aside)."""
    import soket.nn as rnn
    import soket.nn.functional as rfun
    import soket.nn.init as rinit
    import soket.optim as ropt
    import soket.transforms as rtf
    import soket.utils.data as rdata
    import soket_b200.api as mine
    import soket_b200.nn.functional as mfun
    import soket_b200.nn.init as minit
    from soket_b200 import nn as mnn, optim as mopt, transforms as mtf
    from soket_b200.utils import data as mdata

    def pub(o):
        return {n for n in dir(o) if not n.startswith('_')}
    leaked = {'Sequence', 'autodiff', 'backend', 'creation', 'cupy', 'cupy_device', 'detached', 'device', 'dtype',
              'np', 'numpy', 'ops', 'tensor', 'util', 'warnings', 'module', 'prototypes', 'Tuple', 'soket',
              'dataset', 'loader', 'datasets', 'Optional', 'Callable', 'List', 'ceil', 'Dataset_', 'ABC',
              'abstractmethod'}
    assert pub(ref_soket) - pub(mine) - leaked == set()
    assert pub(rnn) - pub(mnn) - leaked == set()
    assert pub(ropt) - pub(mopt) - leaked == set()
    assert pub(rinit) - pub(minit) - leaked == set()
    assert pub(rfun) - pub(mfun) - leaked == set()
    assert pub(rtf) - pub(mtf) - leaked - {'Tensor'} == set()
    assert pub(rdata) - pub(mdata) - leaked == set()
    assert pub(ref_soket.Tensor) - pub(mine.Tensor) == set()
    assert pub(rnn.Module) - pub(mnn.Module) == set()
    for cls in ('Identity', 'Linear', 'Sequential', 'Residual', 'ReLU', 'SoftmaxCrossEntropyLoss', 'BatchNorm1d',
                'BatchNorm2d', 'BatchNorm3d', 'LayerNorm', 'Dropout'):
        assert pub(getattr(rnn, cls)) - pub(getattr(mnn, cls)) == set(), cls
    for cls in ('Optimizer', 'SGD', 'Adam'):
        assert pub(getattr(ropt, cls)) - pub(getattr(mopt, cls)) == set(), cls
    assert pub(ref_soket.Device) - pub(mine.Device) == set()
    assert int(mine.DeviceType.GPU) == int(ref_soket.DeviceType.GPU) and mine.gpu().type == mine.DeviceType.GPU
    with mine.lazy():
        pass


def test_gradient_arena_plan_reverse_order_aligned_buckets():
    """soket_b200.dp.plan_arena (host logic of the data-parallel gradient arena, SURVEY.md section 8e:
    "bucketed in reverse layer order"): slots in reverse parameter order, 256-byte aligned,
    non-overlapping, buckets contiguous and covering every parameter exactly once."""
    from soket_b200.dp import SLOT_ALIGN, plan_arena
    sizes = [784 * 4096, 4096] + [4096 * 4096, 4096, 4096, 4096] * 4 + [4096 * 10, 10]
    offsets, total, buckets = plan_arena(sizes, 16 << 20)
    order = sorted(range(len(sizes)), key=lambda i: offsets[i])
    assert order == list(reversed(range(len(sizes))))            # backward reaches the last layer first
    end = 0
    for i in order:
        assert offsets[i] % SLOT_ALIGN == 0 and offsets[i] >= end
        end = offsets[i] + sizes[i]
    assert total >= end and total % SLOT_ALIGN == 0
    seen = []
    pos = 0
    for start, stop, members in buckets:
        assert start == pos and stop > start
        pos = stop
        for i in members:
            assert start <= offsets[i] and offsets[i] + sizes[i] <= stop
        seen += members
    assert pos == total and sorted(seen) == list(range(len(sizes)))
    assert all(stop - start >= (16 << 20) for start, stop, _ in buckets[:-1])
    # degenerate inputs
    assert plan_arena([], 1024) == ([], 0, [])
    assert plan_arena([5], 1 << 30) == ([0], SLOT_ALIGN, [(0, SLOT_ALIGN, [0])])


def test_adam_ratio_bound_is_a_bound():
    """|m_hat / sqrt(v_hat)| of Adam never exceeds adam_ratio_bound, whatever the gradient sequence
    (here: adversarial ones -- constant, alternating, one spike, growing, shrinking)."""
    from soket_b200.optim import adam_ratio_bound
    for b1, b2 in [(0.9, 0.999), (0.5, 0.9), (0.0, 0.99), (0.95, 0.95)]:
        sup = adam_ratio_bound(b1, b2)
        T = 400
        seqs = [np.ones(T), (-1.0) ** np.arange(T), np.r_[np.zeros(T - 1) + 1e-8, 1.0], 1.1 ** np.arange(T),
                0.9 ** np.arange(T), np.r_[1.0, np.zeros(T - 1) + 1e-12]]
        # the maximiser of the Cauchy-Schwarz step: g_{t-k} proportional to (b1 / b2)^k
        seqs.append((b1 / b2) ** np.arange(T)[::-1] if b1 > 0 else np.ones(T))
        for g in seqs:
            m = v = 0.0
            for t, gt in enumerate(g, 1):
                m = b1 * m + (1 - b1) * gt
                v = b2 * v + (1 - b2) * gt * gt
                r = abs(m / (1 - b1 ** t)) / np.sqrt(v / (1 - b2 ** t))
                assert r <= adam_ratio_bound(b1, b2, t) * (1 + 1e-9), (b1, b2, t, r)
                if sup is not None:
                    assert r <= sup * (1 + 1e-9)
    assert adam_ratio_bound(0.9, 1.0) is None and adam_ratio_bound(0.999, 0.9) is None


def test_peer_memory_shards_cover_every_element_once():
    """soket_b200.dp.piece_of / shard_len / plan_arena(max_members): host logic of the peer-memory data-parallel
    update (csrc/dp_p2p.cu): a bucket's arena range is cut into `world` contiguous pieces of a multiple of 64
    elements; every element of every tensor falls into exactly one rank's piece; a bucket never holds more
    tensors than one launch takes."""
    from soket_b200.dp import piece_of, plan_arena, shard_len
    sizes = [784 * 4096, 4096, 4096 * 4096, 4096, 4096 * 10, 10, 7, 12]
    for world in (2, 3, 4, 8):
        for bucket_floats in (1 << 10, 1 << 22, 1 << 30):
            offsets, total, buckets = plan_arena(sizes, bucket_floats, max_members=32)
            for start, end, members in buckets:
                L = shard_len(end - start, world)
                assert L % 64 == 0 and L * world >= end - start and L * (world - 1) < end - start + 64 * world
                for i in members:
                    cover = np.zeros(sizes[i], np.int32)
                    for r in range(world):
                        s0, c = piece_of(start, end - start, offsets[i], sizes[i], r, world)
                        assert 0 <= s0 and s0 + c <= sizes[i]
                        if c:
                            lo = start + r * L
                            assert lo <= offsets[i] + s0 and offsets[i] + s0 + c <= min(lo + L, end)
                            assert (offsets[i] + s0) % 64 == 0
                        cover[s0:s0 + c] += 1
                    assert (cover == 1).all()
    offsets, total, buckets = plan_arena([10] * 100, 1 << 30, max_members=32)
    assert [len(m) for _, _, m in buckets] == [32, 32, 32, 4]
    assert sorted(i for _, _, m in buckets for i in m) == list(range(100))
