mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29593 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras > gpurun_out/r2ap_bench_2gpu.json 2> gpurun_out/r2ap_bench_2gpu.err
echo "stdout lines: $(wc -l < gpurun_out/r2ap_bench_2gpu.json)"; head -c 120 gpurun_out/r2ap_bench_2gpu.json; echo
grep -c "NCCL version" gpurun_out/r2ap_bench_2gpu.err
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | wc -l
