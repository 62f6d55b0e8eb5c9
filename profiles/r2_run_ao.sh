mkdir -p gpurun_out
for a in "1600 4 0" "1280 2 0" "2048 8 0" "1024 1 0"; do timeout 600 python profiles/ring_slab.py $a 2>&1 | grep "phase=1" | tee -a gpurun_out/r2ao_ring_slab.txt; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py tests/test_siblings.py tests/test_gpu_dist.py -q -m gpu -x --tb=short -k "pk or Pk or xpk or XPk or full_size or baseline or slab" 2>&1 | tail -3
