mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 profiles/bin_stage.py 2048 2>&1 | grep -v "Warn\|\*\*\*\|OMP" | tee gpurun_out/r2ag_bin_stage_2gpu.txt
