#!/bin/bash
# One-call GPU validation: changed-area tests first, then the rest of the suite, the bench line,
# the config-1 epoch numbers.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== new/changed tests" ; date
timeout 400 python -m pytest tests/test_data_pipeline.py tests/test_example.py tests/test_graph_gpu.py tests/test_fused_gpu.py \
    -m gpu -q --timeout 240 2>&1 | tail -60 > gpurun_out/pytest_new.log
tail -5 gpurun_out/pytest_new.log
echo "== bench n=1" ; date
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
echo "== c1 bench" ; date
timeout 300 python scripts/c1_bench.py --graph --loader > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 1500 gpurun_out/c1_bench.json
echo "== rest of the suite" ; date
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -n 4 \
    --deselect tests/test_data_pipeline.py --deselect tests/test_example.py --deselect tests/test_graph_gpu.py --deselect tests/test_fused_gpu.py \
    2>&1 | tail -60 > gpurun_out/pytest_rest.log
tail -5 gpurun_out/pytest_rest.log
echo "== reference arm" ; date
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 400 gpurun_out/bench_ref.json
date
