mkdir -p gpurun_out
timeout 600 python profiles/step_host_gap.py 1024 PCS 2>&1 | head -2
for i in 1 2; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2al_bench_$i.json 2> gpurun_out/r2al_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2al_bench_$i.json').read().strip().splitlines()[-1])
print('bench run $i', d['ms_per_step'], d['e2e']['ms_per_step'], d['kernels_ms_per_step'], d['stages'])
PY
done
timeout 600 python profiles/step_host_gap.py 1024 PCS 2>&1 | head -2
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
lscpu | grep -i "model name\|^CPU(s)\|MHz" | head -5
