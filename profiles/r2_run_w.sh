set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "ma_vs_oracle_both or clustered or ma_weights or zeldovich_lattice" 2>&1 | tail -3
timeout 400 python profiles/deposit_ab.py 512 1024 2>&1 | grep "TSC\|PCS" | grep "kernel=2" | tee gpurun_out/r2w_deposit_ab.txt
