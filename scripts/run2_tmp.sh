cd /root/repo
bash scripts/dp_p2p_try.sh gpurun_out/r2u
bash scripts/dp_sweep.sh 2 gpurun_out/r2u "SOKET_B200_DP_MODE=p2p" "-"
