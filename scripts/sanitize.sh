#!/bin/bash
# compute-sanitizer passes over the GPU tests (one B200).  memcheck: the whole suite minus the 50-step wide
# trainings; racecheck / synccheck: the kernels that synchronise through shared memory, mbarriers, clusters
# (staged LayerNorm, CTA-pair tcgen05 GEMM with its tile-scheduler ring, split / Adam / fused elementwise).
# Usage: bash scripts/sanitize.sh OUTDIR
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/sanitize}; mkdir -p "$out"
SEL="tests/test_fused_gpu.py tests/test_presplit_gpu.py tests/test_lazy_gpu.py::test_fused_chain_is_one_launch_and_bit_identical"
GEMM='tests/test_matmul_tc_gpu.py -k (f16x3_all_layouts or test_bf16 or linear_bwd_shared_split or tf32_single_pass or epilogues) and not 8192'
run() {  # tool, timeout, pytest args...
  tool=$1; to=$2; shift 2
  echo "# compute-sanitizer --tool $tool python -m pytest $* -m gpu" > "$out/$tool.txt"
  timeout $to compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest "$@" -m gpu -q -x --timeout 1500 -p no:cacheprovider \
      2>&1 | grep -v "^$" | tail -40 >> "$out/$tool.txt"
  echo "exit: ${PIPESTATUS[0]}" >> "$out/$tool.txt"
  tail -4 "$out/$tool.txt"
}
run memcheck 900 tests --deselect tests/test_wide_path_gpu.py::test_wide_model_50_steps_teacher_forced \
    --deselect tests/test_wide_path_gpu.py::test_wide_model_50_steps_free_running --deselect tests/test_dp_gpu.py
run racecheck 600 $SEL
run synccheck 600 $SEL
tool=racecheck_gemm
echo "# compute-sanitizer --tool racecheck python -m pytest $GEMM -m gpu" > "$out/$tool.txt"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_matmul_tc_gpu.py \
    -k "(f16x3_all_layouts or test_bf16 or linear_bwd_shared_split or tf32_single_pass or epilogues) and not 8192" \
    -m gpu -q -x --timeout 1500 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 >> "$out/$tool.txt"
echo "exit: ${PIPESTATUS[0]}" >> "$out/$tool.txt"; tail -4 "$out/$tool.txt"
tool=synccheck_gemm
echo "# compute-sanitizer --tool synccheck python -m pytest $GEMM -m gpu" > "$out/$tool.txt"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_matmul_tc_gpu.py \
    -k "(f16x3_all_layouts or test_bf16 or linear_bwd_shared_split or tf32_single_pass or epilogues) and not 8192" \
    -m gpu -q -x --timeout 1500 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40 >> "$out/$tool.txt"
echo "exit: ${PIPESTATUS[0]}" >> "$out/$tool.txt"; tail -4 "$out/$tool.txt"
