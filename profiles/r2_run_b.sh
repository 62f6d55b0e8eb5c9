set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ma or MA or masc or cabi" --tb=short > gpurun_out/r2b_pytest_ma.log 2>&1
tail -25 gpurun_out/r2b_pytest_ma.log
timeout 600 python profiles/deposit_ab.py 512 1024 > gpurun_out/r2b_deposit_ab.txt 2>&1
cat gpurun_out/r2b_deposit_ab.txt
timeout 900 python -m pytest tests/test_gpu_baseline_parity.py -q -m gpu --tb=short > gpurun_out/r2b_pytest_baseline.log 2>&1
tail -15 gpurun_out/r2b_pytest_baseline.log
timeout 900 python -m pytest tests -q -m gpu --tb=short --deselect tests/test_gpu_baseline_parity.py > gpurun_out/r2b_pytest_all.log 2>&1
tail -8 gpurun_out/r2b_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_bench.err; cut -c1-1500 gpurun_out/r2b_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -3 gpurun_out/r2b_smoke.log
