#!/bin/bash
# Round-end measurement series on one B200: full GPU suite, smoke, bench line, config-1 epoch numbers,
# ncu launch list of the bench command.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tests"; date
timeout 500 python -m pytest tests -m gpu -q --timeout 300 -n 4 2>&1 | tail -15 > gpurun_out/pytest_final.log; tail -3 gpurun_out/pytest_final.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; date
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -c 300 gpurun_out/bench_final_n1.json
echo "== c1"; date
timeout 200 python scripts/c1_bench.py --graph --loader > gpurun_out/c1_final.json 2> gpurun_out/c1_final.err; tail -c 200 gpurun_out/c1_final.json
echo "== ncu launch list"; date
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 300 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches_final.csv; date
echo "== ncu --set full of the hot kernels"; date
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"gemm_tc2_kernel|ln_fwd_staged|ln_bwd_staged|adam_kernel|split_rows|split_lo|ln_param_grads|colsum_partials" -c 24 \
    -o gpurun_out/prof_final python scripts/profile_kernels.py 1 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
ls -la gpurun_out/prof_final* ; date
echo "== LayerNorm kernels standalone"; timeout 100 python scripts/ln_bench.py 2>&1 | tail -6 | tee gpurun_out/ln_bench_final.txt
