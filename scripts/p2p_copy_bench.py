"""Copy-engine bandwidth over NVLink between two ranks' IPC-mapped buffers (launch under torchrun, >= 2 ranks):
rank 0 pulls from / pushes to rank 1's buffer with cudaMemcpyAsync, 1 / 2 / 4 streams, several sizes."""
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import soket_b200 as sk                                   # noqa: E402
from soket_b200 import _core as B, _fused as F, dp        # noqa: E402


def main():
    env = dp.read_env()
    sk.init(env.local_rank)
    rdv = dp.Rendezvous(env)
    n = 128 << 20
    buf = B.zeros((n // 4,), "float32")
    local = B.zeros((n // 4,), "float32")
    sk.synchronize()
    handles = [pickle.loads(x) for x in rdv.all_gather_bytes(pickle.dumps(F.ipc_export(buf)))]
    if env.rank == 0:
        peer = F.ipc_open(*handles[1])
        for mb in (1, 4, 8, 16, 64):
            nbytes = mb << 20
            copies = min(8, n // nbytes)
            for streams in (1, 2, 4):
                for name, dst, src in (("pull", local.data_ptr, peer), ("push", peer, local.data_ptr)):
                    F.p2p_copy_probe(dst, src, nbytes, copies, streams, 2)
                    ms = F.p2p_copy_probe(dst, src, nbytes, copies, streams, 10) / 10
                    print(f"{name} {copies} x {mb:3d} MB on {streams} stream(s): {ms * 1e3 / copies:7.1f} us per copy  "
                          f"{copies * nbytes / ms / 1e6:7.1f} GB/s", flush=True)
        ms = F.p2p_copy_probe(local.data_ptr + (64 << 20), local.data_ptr, 64 << 20, 1, 1, 10) / 10
        print(f"local D2D 64 MB: {(64 << 20) / ms / 1e6:7.1f} GB/s (read + write = 2x)", flush=True)
    rdv.barrier()
    if env.rank == 0:
        F.ipc_close_all()
    rdv.barrier()
    rdv.close()


if __name__ == "__main__":
    main()
