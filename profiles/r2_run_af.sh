mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring2_kernel -s 3 -c 1 -o gpurun_out/r2af_ring2_slab python profiles/ring_slab.py 2048 8 0 > gpurun_out/r2af_ncu.log 2>&1
tail -2 gpurun_out/r2af_ncu.log
