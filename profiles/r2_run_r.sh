set -x
mkdir -p gpurun_out
G=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29551 profiles/dist_stages.py 1024 PCS > gpurun_out/r2r_dist_stages_${G}gpu.txt 2>&1
grep -v "Warn\|\*\*\*\|OMP" gpurun_out/r2r_dist_stages_${G}gpu.txt
python -m pytest tests/test_gpu_dist.py -q -m gpu --tb=short 2>&1 | tail -3
