cd /root/repo
mkdir -p gpurun_out/r2q
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29990 scripts/nccl_bench.py > gpurun_out/r2q/nccl_bench.log 2>&1
grep -i "all-reduce\|nvls\|error" gpurun_out/r2q/nccl_bench.log | grep -v "^\[.*TUNING" | head -30
bash scripts/dp_sweep.sh 8 gpurun_out/r2q "SOKET_B200_GEMM_RESERVE_SMS=16" "SOKET_B200_GEMM_RESERVE_SMS=20 SOKET_B200_DP_OPT_OVERLAP=0" "SOKET_B200_GEMM_RESERVE_SMS=32"
