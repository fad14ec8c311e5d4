#!/bin/bash
# bash scripts/sweep_env.sh VAR v1 v2 ... : bench.py (N=1, 6 steps) per value; prints ms/step and the GEMM family
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/sweep_${var}_${v}.json
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_${var}_${v}.json"))
f=d["kernel_families"]
print("$var=$v", "ms/step %.2f" % d["ms_per_step"], "gemm %.2f ms %.1f TF" % (f["gemm_tc"]["ms_per_step"], f["gemm_tc"]["rate"]),
      "prep %.2f" % f["gemm_prep"]["ms_per_step"], "ln %.2f+%.2f" % (f["ln_fwd"]["ms_per_step"], f["ln_bwd"]["ms_per_step"]))
PY
done
