#!/bin/bash
# bash scripts/dp_sweep.sh N OUTDIR "ENV1=a ENV2=b" "ENV3=c" ... : one strong-scaling bench.py run at N GPUs per
# environment variant ("-" = none); prints ms/step, the kernel families and the DP phase timeline of each.
cd "$(dirname "$0")/.."
N=$1; out=$2; shift 2
mkdir -p "$out"
i=0
for v in "$@"; do
  i=$((i + 1))
  tag=$(echo "$v" | tr ' =/' '___')
  [ "$v" = "-" ] && v="" && tag=base
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $((29700 + i)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline \
      > "$out/sweep_n${N}_${tag}.json" 2> "$out/sweep_n${N}_${tag}.err"
  python - "$out/sweep_n${N}_${tag}.json" "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    f = d["kernel_families"]
    t = d.get("dp_timeline", {})
    print(f"[{sys.argv[2] or 'base'}] {d['ms_per_step']:.2f} ms/step  " +
          " ".join(f"{k}={v['ms_per_step']:.2f}" for k, v in f.items() if v['ms_per_step'] > 0.05) +
          "  | " + " ".join(f"{k}={v:.2f}" for k, v in t.items() if isinstance(v, float)), flush=True)
except Exception as e:
    print(f"[{sys.argv[2]}] FAILED: {e}", flush=True)
PY
done
