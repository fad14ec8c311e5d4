"""All-reduce microbenchmark through soket_b200's own NCCL binding (launch under torchrun, one rank per GPU):
fp32 sum all-reduce of 12.8 MB .. 1.09 GB on the comm stream of an otherwise idle GPU, CUDA-event timed,
max over ranks.  Tells what the collective costs by itself, next to what bench.py's DP timeline sees beside backward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import soket_b200 as sk                                   # noqa: E402
from soket_b200 import _core as B, _fused as F, dp        # noqa: E402


def main():
    env = dp.read_env()
    sk.init(env.local_rank)
    rdv = dp.Rendezvous(env)
    F.nccl_init(env.rank, env.world, dp.exchange_unique_id(rdv, F.nccl_unique_id))
    total = 272_000_000
    arena = B.zeros((total,), "float32")
    for mb, reps in ((12.8, 20), (67.2, 20), (268.8, 10), (1088.0, 5)):
        n = int(mb * 1e6 / 4)
        view = arena[:n]
        pieces = 1
        for chunks in (1, 16 if mb > 1000 else 1):
            step = n // chunks
            for _ in range(3):
                F.nccl_allreduce_on(view, B.STREAM_COMM)
            sk.synchronize(); rdv.barrier()
            e0, e1 = sk.Event(), sk.Event()
            e0.record(B.STREAM_COMM)
            for _ in range(reps):
                for c in range(chunks):
                    F.nccl_allreduce_on(arena[c * step:(c + 1) * step], B.STREAM_COMM)
            e1.record(B.STREAM_COMM)
            e1.synchronize()
            ms = max(rdv.all_gather_float(e0.elapsed_ms(e1) / reps))
            if env.rank == 0:
                print(f"all-reduce {mb:7.1f} MB in {chunks:2d} call(s): {ms:7.3f} ms  algbw {mb / ms:6.1f} GB/s  "
                      f"busbw {mb / ms * 2 * (env.world - 1) / env.world:6.1f} GB/s", flush=True)
    rdv.barrier()
    sk.synchronize()
    F.nccl_destroy()
    rdv.close()


if __name__ == "__main__":
    main()
