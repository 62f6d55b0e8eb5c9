set -x
mkdir -p gpurun_out
G=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $G --workload cfg4_2048_strong --steps 2 --warmup 3 --no-extras > gpurun_out/r2n_bench_2048_strong_${G}gpu.json 2> gpurun_out/r2n_bench_${G}gpu.err
tail -3 gpurun_out/r2n_bench_${G}gpu.err; cut -c1-700 gpurun_out/r2n_bench_2048_strong_${G}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $G --steps 3 --warmup 3 --no-extras > gpurun_out/r2n_bench_${G}gpu.json 2>> gpurun_out/r2n_bench_${G}gpu.err
tail -3 gpurun_out/r2n_bench_${G}gpu.err; cut -c1-700 gpurun_out/r2n_bench_${G}gpu.json
