#!/bin/bash
# bash scripts/dp_diag.sh OUTDIR : the 2-rank DP parity tests, then per-tensor gradient diagnostics (DP_DIAG=1)
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/dp_diag}; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
python -m pytest tests/test_dp_gpu.py -m gpu -q -rA 2>&1 | tail -60 > $out/pytest_dp_n2.log
tail -8 $out/pytest_dp_n2.log
i=0
for cfg in "2 0 sgd" "2 1 sgd" "2 1 adam"; do
  set -- $cfg; W=$1; wide=$2; opt=$3
  i=$((i+1))
  echo "=== W=$W wide=$wide $opt" | tee -a $out/diag.log
  env DP_DIAG=1 DP_NORM=layer DP_OPT=$opt DP_WIDE=$wide timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W \
     --master-addr 127.0.0.1 --master-port $((29800+i)) scripts/dp_parity.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -40 >> $out/diag.log
done
