#!/bin/bash
# racecheck + synccheck over one staged-LayerNorm case pair and two CTA-pair GEMM cases (a 2-minute slice of
# scripts/sanitize.sh for when the GPU budget does not allow the whole recipe)
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/sanitize_quick}; mkdir -p $out
SEL='tests/test_fused_gpu.py::test_layernorm_fwd_bwd[16-4096] tests/test_fused_gpu.py::test_layernorm_fwd_bwd[50-2048] tests/test_matmul_tc_gpu.py::test_f16x3_all_layouts[False-False-256-384-512] tests/test_matmul_tc_gpu.py::test_f16x3_all_layouts[True-False-512-784-256]'
for tool in racecheck synccheck; do
  echo "# compute-sanitizer --tool $tool python -m pytest $SEL -m gpu -q" > $out/$tool.txt
  timeout 48 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -q -x -p no:cacheprovider 2>&1 | grep -v "^$" | tail -15 >> $out/$tool.txt
  echo "exit: ${PIPESTATUS[0]}" >> $out/$tool.txt
  tail -3 $out/$tool.txt
done
