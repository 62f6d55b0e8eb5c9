mkdir -p gpurun_out
G=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus $G --steps 3 --warmup 3 > gpurun_out/r2ah_bench_${G}gpu.json 2> gpurun_out/r2ah_bench_${G}gpu.err
tail -2 gpurun_out/r2ah_bench_${G}gpu.err | cut -c1-300; cut -c1-700 gpurun_out/r2ah_bench_${G}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29582 profiles/dist_stages.py 1024 PCS 2>&1 | grep -v "Warn\|\*\*\*\|OMP" | tee gpurun_out/r2ah_dist_stages_${G}gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29583 profiles/bin_stage.py 2048 2>&1 | grep -v "Warn\|\*\*\*\|OMP" | tee gpurun_out/r2ah_bin_stage_${G}gpu.txt
