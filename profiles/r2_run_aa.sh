mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aa_launches_1024_pcs.csv python profiles/ncu_step.py 1024 PCS 2 > gpurun_out/r2aa_ncu.log 2>&1
tail -2 gpurun_out/r2aa_ncu.log
