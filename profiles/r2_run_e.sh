set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "vs_oracle or clustered or weights or deep or zeldovich or golden" --tb=short > gpurun_out/r2e_pytest_ma.log 2>&1
tail -12 gpurun_out/r2e_pytest_ma.log
timeout 600 python profiles/deposit_ab.py 512 1024 > gpurun_out/r2e_deposit_ab.txt 2>&1
cat gpurun_out/r2e_deposit_ab.txt
timeout 900 python -m pytest tests/test_gpu_baseline_parity.py -q -m gpu --tb=short > gpurun_out/r2e_pytest_baseline.log 2>&1
tail -12 gpurun_out/r2e_pytest_baseline.log
