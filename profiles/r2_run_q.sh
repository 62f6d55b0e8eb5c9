set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2q_pytest_all.log 2>&1
tail -6 gpurun_out/r2q_pytest_all.log
timeout 900 python bench.py > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
tail -3 gpurun_out/r2q_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['check']['parity'])
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('kernels_ms_per_step'))
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
