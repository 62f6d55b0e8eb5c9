set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 --no-extras > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err
tail -3 gpurun_out/r2j_bench_8gpu.err; cat gpurun_out/r2j_bench_8gpu.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 profiles/dist_stages.py 1024 PCS > gpurun_out/r2j_dist_stages_8gpu.txt 2>&1
grep -v "Warn\|\*\*\*\|OMP" gpurun_out/r2j_dist_stages_8gpu.txt
