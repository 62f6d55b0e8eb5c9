mkdir -p gpurun_out
timeout 600 python profiles/ring_slab.py 2048 8 0 2>&1 | tee gpurun_out/r2ad_ring_slab.txt
timeout 600 python profiles/ring_slab.py 1024 1 0 2>&1 | tee -a gpurun_out/r2ad_ring_slab.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py tests/test_siblings.py tests/test_gpu_dist.py -q -m gpu -x --tb=short -k "pk or Pk or xpk or XPk or full_size or baseline or slab" 2>&1 | tail -4
