set -x
mkdir -p gpurun_out
# launch list of one warm step of the default workload (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/r2k_launches_1024_pcs.csv python profiles/ncu_step.py 1024 PCS 2 > gpurun_out/r2k_ncu.log 2>&1
# full captures of the kernels DESIGN.md names (512^3 so that ~40 replays stay short), one launch each
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_lane -s 1 -c 1 -o gpurun_out/r2k_lane_pcs python profiles/ncu_deposit.py 512 PCS 0 >> gpurun_out/r2k_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_tile -s 1 -c 1 -o gpurun_out/r2k_tile_cic python profiles/ncu_deposit.py 512 CIC 0 >> gpurun_out/r2k_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bin_pass -s 2 -c 2 -o gpurun_out/r2k_pass python profiles/ncu_deposit.py 512 CIC 0 >> gpurun_out/r2k_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring2_kernel -s 1 -c 1 -o gpurun_out/r2k_ring2_1024 python profiles/ncu_step.py 1024 CIC 2 >> gpurun_out/r2k_ncu.log 2>&1
tail -4 gpurun_out/r2k_ncu.log
ls -la gpurun_out/r2k_*
# strong-scaling workload: the small version of the code path, then BASELINE configs[3] on one GPU (streamed batches)
timeout 600 python bench.py --workload 512_strong --steps 3 --warmup 3 --no-extras > gpurun_out/r2k_bench_512_strong.json 2> gpurun_out/r2k_bench.err
cut -c1-600 gpurun_out/r2k_bench_512_strong.json
timeout 900 python bench.py --workload cfg4_2048_strong --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2k_bench_2048_strong_1gpu.json 2>> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_bench.err; cut -c1-900 gpurun_out/r2k_bench_2048_strong_1gpu.json
