mkdir -p gpurun_out
G=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $G --steps 3 --warmup 3 --no-extras > gpurun_out/r2am_bench_${G}gpu.json 2> gpurun_out/r2am_bench_${G}gpu.err
tail -2 gpurun_out/r2am_bench_${G}gpu.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2am_bench_${G}gpu.json').read().strip().splitlines()[-1])
print($G, 'gpus', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['ring_kernel_only']['frac'], d['check']['parity']['ok'], d['config']['grid'])
PY
