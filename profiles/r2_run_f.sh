set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "vs_oracle or clustered" --tb=line > gpurun_out/r2f_pytest_ma.log 2>&1
tail -12 gpurun_out/r2f_pytest_ma.log
timeout 600 python profiles/deposit_ab.py 512 > gpurun_out/r2f_deposit_ab.txt 2>&1
cat gpurun_out/r2f_deposit_ab.txt
