set -x
mkdir -p gpurun_out
timeout 600 python profiles/deposit_ab.py 512 > gpurun_out/r2g_deposit_ab.txt 2>&1
grep PCS gpurun_out/r2g_deposit_ab.txt
