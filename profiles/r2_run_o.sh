set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "xpk or ring2 or pk_golden or pk_vs_oracle or vs_oracle_both or full_size_1024" --tb=short 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_siblings.py tests/test_drivers.py -q -m gpu --tb=short 2>&1 | tail -4
timeout 900 python bench.py --workload cfg5_1024_xpk --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2o_bench_xpk.json 2> gpurun_out/r2o_bench.err
tail -3 gpurun_out/r2o_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_bench_xpk.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline'], d['kernels_ms_per_step'])
PY
