mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ab_launches_pk_1024.csv python profiles/ncu_pk.py 1024 2 > gpurun_out/r2ab_ncu.log 2>&1
tail -2 gpurun_out/r2ab_ncu.log
