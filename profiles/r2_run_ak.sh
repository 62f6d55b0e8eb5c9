mkdir -p gpurun_out
timeout 600 python profiles/step_host_gap.py 1024 PCS 2>&1 | tee gpurun_out/r2ak_step_host_gap.txt
