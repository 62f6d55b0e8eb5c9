set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dist_stages.py 1024 PCS > gpurun_out/r2i_dist_stages_2gpu.txt 2>&1
cat gpurun_out/r2i_dist_stages_2gpu.txt | grep -v Warn
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err
tail -3 gpurun_out/r2i_bench_2gpu.err; cat gpurun_out/r2i_bench_2gpu.json | cut -c1-3000
python -m pytest tests/test_gpu_dist.py -q -m gpu --tb=short 2>&1 | tail -5
