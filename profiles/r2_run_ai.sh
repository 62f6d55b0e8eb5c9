mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "ma_vs_oracle_both or clustered or ma_weights or zeldovich_lattice or deep_sort or host_chunked" 2>&1 | tail -3
timeout 400 python profiles/deposit_ab.py 1024 2>&1 | grep "kernel=2\|CIC\|NGP" | tee gpurun_out/r2ai_deposit_ab.txt
