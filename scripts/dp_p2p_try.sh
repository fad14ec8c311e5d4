#!/bin/bash
# bash scripts/dp_p2p_try.sh OUTDIR : the peer-memory DP mode at 2 ranks, one parity run per configuration
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/p2p}; mkdir -p $out
i=0
for cfg in "layer adam 0" "layer adam 1" "batch adam 0"; do
  set -- $cfg; i=$((i+1))
  echo "=== p2p norm=$1 opt=$2 wide=$3" | tee -a $out/p2p.log
  env DP_MODE=p2p DP_NORM=$1 DP_OPT=$2 DP_WIDE=$3 timeout -k 10 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
     --master-addr 127.0.0.1 --master-port $((29850+i)) scripts/dp_parity.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -30 >> $out/p2p.log
  tail -3 $out/p2p.log
done
