mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ae_launches_ring_slab.csv python profiles/ring_slab.py 2048 8 0 > gpurun_out/r2ae_ncu.log 2>&1
tail -2 gpurun_out/r2ae_ncu.log
