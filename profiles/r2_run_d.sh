set -x
mkdir -p gpurun_out
./profiles/microbench/atoms_pattern > gpurun_out/r2d_atoms_pattern.txt 2>&1
cat gpurun_out/r2d_atoms_pattern.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "vs_oracle or clustered or weights or deep or zeldovich" --tb=short > gpurun_out/r2d_pytest_ma.log 2>&1
tail -5 gpurun_out/r2d_pytest_ma.log
timeout 600 python profiles/deposit_ab.py 512 1024 > gpurun_out/r2d_deposit_ab.txt 2>&1
cat gpurun_out/r2d_deposit_ab.txt
