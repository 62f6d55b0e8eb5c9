mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_parity.py tests/test_siblings.py -q -m gpu -x --tb=short -k "pk or Pk or xpk or XPk or full_size or baseline" 2>&1 | tail -4
timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench.err
tail -3 gpurun_out/r2z_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['roofline'])[:700])
PY
