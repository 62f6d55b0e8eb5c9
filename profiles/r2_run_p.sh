set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r2p_bench_2gpu.json 2> gpurun_out/r2p_bench_2gpu.err
tail -3 gpurun_out/r2p_bench_2gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e'], d['check']['parity']['ok'], d['kernels_ms_per_step'])
PY
python -m pytest tests/test_gpu_dist.py -q -m gpu --tb=short 2>&1 | tail -3
