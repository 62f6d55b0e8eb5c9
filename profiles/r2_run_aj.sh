mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r2aj_pytest_all.log 2>&1
tail -5 gpurun_out/r2aj_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2aj_bench_1gpu.json 2> gpurun_out/r2aj_bench.err
tail -3 gpurun_out/r2aj_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aj_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['roofline']['frac'], d['check']['parity'])
for k,v in d['extra'].items(): print(k, v.get('ms_per_step'), v.get('kernels_ms_per_step'))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2aj_bench_1gpu_reference.json 2>> gpurun_out/r2aj_bench.err
cut -c1-400 gpurun_out/r2aj_bench_1gpu_reference.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_lane -s 1 -c 1 -o gpurun_out/r2aj_lane_pcs python profiles/ncu_deposit.py 512 PCS 0 > gpurun_out/r2aj_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_lane -s 1 -c 1 -o gpurun_out/r2aj_lane_tsc python profiles/ncu_deposit.py 512 TSC 0 >> gpurun_out/r2aj_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2aj_launches_1024_pcs.csv python profiles/ncu_step.py 1024 PCS 2 >> gpurun_out/r2aj_ncu.log 2>&1
tail -2 gpurun_out/r2aj_ncu.log
