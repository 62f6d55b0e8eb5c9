mkdir -p gpurun_out
timeout 600 python profiles/ring_slab.py 2048 8 0 2>&1 | tee gpurun_out/r2ac_ring_slab.txt
timeout 600 python profiles/ring_slab.py 2048 8 3 2>&1 | tee -a gpurun_out/r2ac_ring_slab.txt
timeout 600 python profiles/ring_slab.py 1024 1 0 2>&1 | tee -a gpurun_out/r2ac_ring_slab.txt
