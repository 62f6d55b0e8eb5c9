set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "vs_oracle or clustered or weights or deep or zeldovich or golden" --tb=short 2>&1 | tail -6
timeout 600 python profiles/deposit_ab.py 512 > gpurun_out/r2m_deposit_ab.txt 2>&1
cat gpurun_out/r2m_deposit_ab.txt
