"""Data-parallel parity worker (launch under torchrun, one rank per GPU).

Each rank trains the same small MLPResNet on ITS row shard of a seeded global batch for a
few Adam steps with the NCCL gradient all-reduce of soket_b200.dp; rank 0 then compares
the parameters with the oracle's W-shard emulation (W forward/backwards on
the shards, gradients averaged, one optimiser step).  Exit code 0 = parity holds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import soket_b200 as sk                                   # noqa: E402
import soket_b200.api as soket                            # noqa: E402
from soket_b200 import dp, nn                             # noqa: E402
from soket_b200.optim import SGD, Adam                    # noqa: E402
from oracle import ref_model, soket_np as O               # noqa: E402   (checker only)


def main():
    env = dp.read_env()
    sk.init(env.local_rank)
    rdv = dp.Rendezvous(env)
    # DP_WIDE=1: wide enough for the tcgen05 GEMMs / staged LayerNorm kernels of the benchmark
    wide = os.environ.get("DP_WIDE", "0") == "1"
    dim, hidden, nb, C, B = (784, 512, 3, 10, 1024) if wide else (784, 256, 2, 10, 512)
    B = int(os.environ.get("DP_BATCH", B))     # global rows (>= 256 per rank keeps every Linear on the pre-split GEMM path)
    norm = os.environ.get("DP_NORM", "layer")
    # SGD is linear in the gradient: strict multi-step comparison.  Adam's first steps are
    # lr * g / (|g| + eps): elements whose true gradient is zero (every bias in front of a
    # normalisation layer) turn rounding noise into +-lr updates on BOTH backends, so Adam is
    # compared after one step on the elements whose gradient is above the noise floor.
    use_adam = os.environ.get("DP_OPT", "sgd") == "adam"
    # BatchNorm at this batch size amplifies rounding-level differences chaotically from the
    # second step on (tests/test_dropin_gpu.py): one exact step there
    steps = 1 if (use_adam or norm == "batch") else 3
    rng = np.random.default_rng(int(os.environ.get("DP_SEED", "0")))
    om = O.MLPResNet(dim, hidden, nb, C, norm=norm)
    om.init_kaiming(0)
    for k in om.params:
        if k.endswith(".W"):
            om.params[k] = (rng.standard_normal(om.params[k].shape) * np.sqrt(2.0 / om.params[k].shape[0])).astype("float32")
    model = ref_model.build_model(nn, dim, hidden, nb, C, norm=norm, drop_prob=0.0)
    named = ref_model.named_parameters(model, nb)
    for k, t in named.items():
        # rank 0 holds the real weights, the others garbage: broadcast_parameters must fix that
        w = om.params[k] if env.rank == 0 else np.full_like(om.params[k], 7.0)
        t.data = soket.Tensor(w.copy())
    opt = Adam(model.parameters(), lr=1e-3) if use_adam else SGD(model.parameters(), lr=0.05)
    # DP_MODE=p2p: reduce-scatter + Adam + operand split + all-gather as one kernel per bucket over NVLink peer
    # memory instead of ncclAllReduce + replicated Adam; share_grads leaves the reduced gradient in every
    # replica's arena so that the gradient check below reads the same thing in both modes
    mode = os.environ.get("DP_MODE", "nccl")
    ddp = dp.DataParallel(opt, rdv, bucket_mb=float(os.environ.get("DP_BUCKET_MB", "0.25")), mode=mode,
                          share_grads=True, lazy_master=os.environ.get("DP_LAZY", "0") == "1")   # several buckets
    ddp.broadcast_parameters(0)
    crit = nn.SoftmaxCrossEntropyLoss()
    oo = O.Adam(len(om.names()), lr=1e-3) if use_adam else O.SGD(len(om.names()), lr=0.05)
    start = {k: v.copy() for k, v in om.params.items()}
    signal = {}
    sl = dp.shard_rows(B, env.rank, env.world)
    losses = []
    total_flips = 0
    for step in range(steps):
        X = rng.random((B, dim), dtype=np.float32)
        y = rng.integers(0, C, B).astype(np.uint8)
        # every rank's ReLU sign pattern of its shard (read off the fused forward kernels' outputs): of the
        # ~1e6 ReLU inputs of a step one may land within rounding distance of zero and get a different 0/1
        # derivative on the two backends, which moves O(1/rows) of every upstream gradient; the oracle's
        # backward of shard r uses rank r's pattern after checking it differs only at such entries
        # (tests/test_wide_path_gpu.py does the same on one GPU)
        signs = ref_model.device_relu_signs(model, X[sl], nb)
        packed = rdv.all_gather_bytes(b"".join(np.packbits(m).tobytes() for m in signs))
        loss = crit(model(soket.Tensor(X[sl])), soket.Tensor(y[sl]))
        loss.backward()
        ddp.step()               # bucketed all-reduce + per-bucket optimizer update, joined here
        losses.append(loss.item())
        dev_grads = {k: named[k].grad.numpy() for k in om.names()} if (env.rank == 0 and step == 0) else None
        if env.rank == 0:
            # oracle: W shards, averaged gradients, one step
            acc = None
            local = {}
            shard_losses = []
            for r in range(env.world):
                s = dp.shard_rows(B, r, env.world)
                rows = s.stop - s.start
                shard_losses.append(om.loss(om.forward(X[s]), y[s]))
                flat = np.frombuffer(packed[r], np.uint8)
                dev_signs, o = [], 0
                for m in ref_model.oracle_relu_signs(om, nb):
                    nbytes = (m.size + 7) // 8
                    dev_signs.append(np.unpackbits(flat[o:o + nbytes], count=m.size).reshape(m.shape).astype(bool))
                    o += nbytes
                assert rows == dev_signs[0].shape[0] and o == flat.size
                masks, flips = ref_model.reconcile_relu_masks(om, dev_signs, nb)
                total_flips += flips
                assert flips <= 16, flips
                g = om.backward(masks)
                acc = {k: np.asarray(v, dtype=np.float32) / np.float32(env.world) if acc is None
                       else acc[k] + np.asarray(v, dtype=np.float32) / np.float32(env.world) for k, v in g.items()}
                # the local rounding term of each weight gradient dW = in.T @ adj (a sum of `rows` products of
                # either sign, rounded at the size of the terms): the kernel tests' fp32 GEMM bound
                for lay, adj in om.adj_mm.items():
                    xin = (om.tape["X"] if lay == "lin0" else om.tape["out.in"] if lay == "out"
                           else om.tape[lay.split(".")[0]]["in" if lay.endswith("lin1") else "lin2.in"])
                    t = 1e-5 * (np.abs(xin).astype(np.float64).T @ np.abs(adj).astype(np.float64)) / env.world
                    local[lay + ".W"] = t if lay + ".W" not in local else local[lay + ".W"] + t
            names = om.names()
            gscale = max(float(np.abs(acc[k]).max()) for k in names)
            if dev_grads is not None:
                # after step() a parameter's .grad is the SUM over ranks (1/W lives in the optimizer
                # kernel): element-wise against the oracle's averaged shard gradients, |err| <= propagated
                # (2e-5 |want| + 4e-5 rms: the adjoint reaching a layer is a composite of fp32 op results each
                # at 1e-5 of its own scale) + local GEMM term, and 1e-5 per tensor in the rms sense
                for k in names:
                    got = dev_grads[k].astype(np.float64) / env.world
                    want = acc[k].astype(np.float64).reshape(got.shape)
                    rms = float(np.sqrt(np.mean(want * want)))
                    bound = 2e-5 * np.abs(want) + 4e-5 * rms + 1e-7 * gscale
                    if k in local:
                        bound = bound + local[k].reshape(got.shape)
                    ex = np.abs(got - want) / bound
                    rr = float(np.sqrt(np.mean((got - want) ** 2)) / max(rms, 1e-30))
                    if os.environ.get("DP_DIAG"):
                        print(f"diag {k:14s} excess {ex.max():9.2f}  n>1 {int((ex > 1).sum()):7d}/{ex.size}  rms rel {rr:.2e}  rms {rms:.2e}",
                              flush=True)
                        continue
                    assert ex.max() <= 1.0, (k, float(ex.max()))
                    if float(np.abs(want).max()) > 1e-4 * gscale:
                        assert rr <= 1e-5, (k, rr)
            signal = {k: np.abs(acc[k]) > 1e-3 * gscale for k in names}   # clear of the rounding-residue floor
            for k, v in zip(names, oo.step([om.params[k] for k in names], [acc[k] for k in names])):
                om.params[k] = v
            assert abs(losses[-1] - shard_losses[0]) <= 1e-4 * max(1.0, abs(shard_losses[0])), (step, losses[-1], shard_losses[0])
    ddp.sync_parameters()        # lazy_master: the replicas' fp32 copies of the GEMM weights
    if env.rank == 0:
        worst = 0.0
        per = []
        for k in om.names():
            got, want = named[k].numpy(), om.params[k]
            keep = signal[k].reshape(want.shape)
            if not keep.any():
                continue
            # error of the UPDATE relative to the largest update of that tensor
            scale = max(float(np.abs(want - start[k]).max()), 1e-30)
            e = float(np.abs(got - want)[keep].max()) / scale
            per.append((e, k))
            worst = max(worst, e)
        print("worst tensors:", sorted(per, reverse=True)[:4], flush=True)
        print(f"dp parity W={env.world} mode={mode} norm={norm} opt={'adam' if use_adam else 'sgd'}: "
              f"worst update rel err {worst:.2e} over {steps} step(s), {total_flips} reconciled ReLU sign(s)", flush=True)
        # a plumbing bug (a tensor not all-reduced, a missing 1/W, a wrong shard) shows up as an O(1)
        # update error; with the ReLU sign patterns reconciled both sides stay at rounding distance
        # (measured 2-rank runs: 0.7e-5 .. 1.5e-5, profiles/r2_dp_parity_n2.log)
        assert worst <= 2e-4, worst
    rdv.barrier()
    ddp.close()
    if use_adam and wide and mode == "nccl":
        fused_split_check(env, rdv, start, dim, hidden, nb, C, B)
    if use_adam and mode == "p2p":
        p2p_matches_nccl_check(env, rdv, start, dim, hidden, nb, C, B)
    rdv.barrier()
    rdv.close()


def p2p_matches_nccl_check(env, rdv, start, dim, hidden, nb, C, B):
    """Six free-running Adam steps in both data-parallel modes from the same start.  The two modes do the same
    arithmetic per element (sum of the ranks' gradients, then the optimizer's update, csrc/optim.cuh) and differ
    in the ORDER of the W-term gradient sum only: at W = 2 (one commutative add) the parameters must be
    bit-identical on every rank, beyond that they agree to rounding.  Also checks that the replicas stay
    replicas and that only 1/W of the optimizer state lives on each rank."""
    runs = {}
    for mode in ("p2p", "nccl"):
        model = ref_model.build_model(nn, dim, hidden, nb, C, norm="layer", drop_prob=0.0)
        named = ref_model.named_parameters(model, nb)
        for k, t in named.items():
            t.data = soket.Tensor((start[k] if env.rank == 0 else np.full_like(start[k], 3.0)).copy())
        opt = Adam(model.parameters(), lr=1e-3)
        ddp = dp.DataParallel(opt, rdv, bucket_mb=0.25, mode=mode, lazy_master=os.environ.get("DP_LAZY", "0") == "1")
        ddp.broadcast_parameters(0)
        crit = nn.SoftmaxCrossEntropyLoss()
        rng = np.random.default_rng(11)
        sl = dp.shard_rows(B, env.rank, env.world)
        losses = []
        for step in range(6):
            X = rng.random((B, dim), dtype=np.float32)
            y = rng.integers(0, C, B).astype(np.uint8)
            loss = crit(model(soket.Tensor(X[sl])), soket.Tensor(y[sl]))
            loss.backward()
            ddp.step()
            losses.append(loss.item())
        ddp.sync_parameters()
        params = {k: t.numpy() for k, t in named.items()}
        digest = repr(float(sum(float(np.abs(v.astype(np.float64)).sum()) for v in params.values())))
        assert len(set(rdv.all_gather_str(digest))) == 1, "replicas diverged in mode " + mode
        if mode == "p2p":
            state = sum(int(m.size) for m in ddp._m if m is not None)
            total = sum(int(t.size) for t in named.values())
            assert state <= total // env.world + 4096, (state, total)
        runs[mode] = (params, losses)
        rdv.barrier()
        ddp.close()
    (pp, lp), (pn, ln) = runs["p2p"], runs["nccl"]
    worst = 0.0
    for k in pp:
        if env.world == 2:
            assert np.array_equal(pp[k], pn[k]), (k, float(np.abs(pp[k] - pn[k]).max()))
        else:
            upd = max(float(np.abs(pn[k] - start[k]).max()), 1e-30)
            worst = max(worst, float(np.abs(pp[k] - pn[k]).max()) / upd)
    if env.world == 2:
        assert lp == ln, (lp, ln)
    else:
        assert worst <= 2e-2, worst         # Adam's lr * sign(g) on elements whose gradient is rounding residue
        assert max(abs(a - b) for a, b in zip(lp, ln)) <= 1e-4, (lp, ln)
    if env.rank == 0:
        print(f"dp p2p vs nccl W={env.world}: 6 Adam steps " + ("bit-identical parameters and losses" if env.world == 2
              else f"worst update difference {worst:.2e}") + f" (losses {lp[0]:.6f} -> {lp[-1]:.6f})", flush=True)


def fused_split_check(env, rdv, start, dim, hidden, nb, C, B):
    """Under data parallelism the per-bucket Adam launches run on the optimizer stream and, from the
    third step on, rewrite each Linear weight's fp16x3 operand split in place (sk_adam_step_split)
    while backward is still running.  Five free-running steps with that fusion on and off must leave
    bit-identical parameters on every rank (same update arithmetic, power-of-two scales only move
    exponents, a 2-rank all-reduce is one commutative add)."""
    from soket_b200 import optim as OPT
    runs = {}
    for fused in (True, False):
        OPT.set_fused_weight_split(fused)
        try:
            model = ref_model.build_model(nn, dim, hidden, nb, C, norm="layer", drop_prob=0.0)
            named = ref_model.named_parameters(model, nb)
            for k, t in named.items():
                t.data = soket.Tensor((start[k] if env.rank == 0 else np.full_like(start[k], 3.0)).copy())
            opt = Adam(model.parameters(), lr=1e-3)
            ddp = dp.DataParallel(opt, rdv, bucket_mb=0.25)
            ddp.broadcast_parameters(0)
            crit = nn.SoftmaxCrossEntropyLoss()
            rng = np.random.default_rng(7)
            sl = dp.shard_rows(B, env.rank, env.world)
            losses, resplits = [], 0
            for step in range(5):
                X = rng.random((B, dim), dtype=np.float32)
                y = rng.integers(0, C, B).astype(np.uint8)
                loss = crit(model(soket.Tensor(X[sl])), soket.Tensor(y[sl]))
                loss.backward()
                ddp.step()
                losses.append(loss.item())
                if step >= 2:
                    n0 = sk.launch_count()
                    sk.get_split(named["blk0.lin1.W"]._data)
                    resplits += sk.launch_count() - n0
            runs[fused] = ({k: t.numpy() for k, t in named.items()}, losses, resplits)
            rdv.barrier()
            ddp.close()
        finally:
            OPT.set_fused_weight_split(True)
    (pf, lf, nf), (pp, lp, np_) = runs[True], runs[False]
    assert nf == 0, nf                       # the weight never needed a split pass of its own
    assert lf == lp, (lf, lp)
    for k in pf:
        assert np.array_equal(pf[k], pp[k]), (k, float(np.abs(pf[k] - pp[k]).max()))
    digest = repr(float(sum(float(np.abs(v).sum()) for v in pf.values())))
    assert len(set(rdv.all_gather_str(digest))) == 1           # and the replicas stayed replicas
    if env.rank == 0:
        print(f"dp fused weight split W={env.world}: 5 Adam steps bit-identical with the fusion on and off "
              f"(losses {lf[0]:.6f} -> {lf[-1]:.6f})", flush=True)


if __name__ == "__main__":
    main()
