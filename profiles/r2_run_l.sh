set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "vs_oracle or chunked or deep or full_size" --tb=short 2>&1 | tail -4
timeout 900 python bench.py --no-extras > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -3 gpurun_out/r2l_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['stages'])
PY
timeout 900 python bench.py --workload cfg4_2048_strong --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_bench_2048_strong_1gpu.json 2>> gpurun_out/r2l_bench.err
tail -3 gpurun_out/r2l_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_2048_strong_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['stages'])
PY
