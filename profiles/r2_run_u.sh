set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "ma_vs_oracle_both or clustered or ma_weights or zeldovich_lattice or deep_sort or host" 2>&1 | tail -5
timeout 400 python profiles/deposit_ab.py 512 1024 2>&1 | grep "TSC\|PCS" | tee gpurun_out/r2u_deposit_ab.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench.err
tail -3 gpurun_out/r2u_bench.err; cut -c1-1500 gpurun_out/r2u_bench_1gpu.json
