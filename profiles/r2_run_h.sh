set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2h_pytest_all.log 2>&1
tail -15 gpurun_out/r2h_pytest_all.log
timeout 600 python profiles/deposit_ab.py 512 1024 > gpurun_out/r2h_deposit_ab.txt 2>&1
cat gpurun_out/r2h_deposit_ab.txt
timeout 900 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -3 gpurun_out/r2h_bench.err; cut -c1-400 gpurun_out/r2h_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r2h_bench_reference.json 2>> gpurun_out/r2h_bench.err
cut -c1-300 gpurun_out/r2h_bench_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
