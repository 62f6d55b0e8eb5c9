set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_lane -s 1 -c 1 -o gpurun_out/r2v_lane_pcs python profiles/ncu_deposit.py 512 PCS 0 > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log
