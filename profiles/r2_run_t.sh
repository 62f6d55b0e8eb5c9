set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_col -s 1 -c 1 -o gpurun_out/r2t_col_pcs python profiles/ncu_deposit.py 512 PCS 3 > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log
ls -la gpurun_out/r2t_*
