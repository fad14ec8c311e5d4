#!/bin/bash
# bash scripts/smallbatch.sh OUTDIR : one-GPU bench at the per-GPU batch sizes of the strong-scaling run (8192/W rows)
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/smallbatch}; mkdir -p $out
for b in 1024 128 2048; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --global-batch $b > $out/bench_b$b.json 2> $out/bench_b$b.err
  tail -1 $out/bench_b$b.err
  python - $out/bench_b$b.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"{d['config']['batch_per_gpu']} rows: {d['ms_per_step']:.2f} ms/step host {d['host_issue_ms_per_step']:.2f} " +
      " ".join(f"{k}={v['ms_per_step']:.2f}" for k, v in d["kernel_families"].items() if v['ms_per_step'] > 0.05))
PY
done
