#!/bin/bash
# bash scripts/gpu_dp2.sh OUTDIR : 2-GPU box -- the real-NCCL / peer-memory DP parity tests, then the strong-scaling bench line
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/dp2}; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -rA 2>&1 | tail -25 > $out/pytest_dp_n2.log
tail -4 $out/pytest_dp_n2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n2.json 2> $out/bench_n2.err
tail -c 400 $out/bench_n2.json
