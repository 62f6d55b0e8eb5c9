# round-1 final single-GPU evidence: bench lines (ours + reference arm), other workloads, ring variants, ncu launch list
set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_uniform.json 2> gpurun_out/bench_r1_uniform.err
python bench.py --steps 10 --warmup 3 --data zeldovich --no-cpu-baseline > gpurun_out/bench_r1_zeldovich.json 2>> gpurun_out/bench_r1_uniform.err
python bench.py --steps 5 --warmup 3 --workload cfg3_1024_tsc --no-cpu-baseline > gpurun_out/bench_r1_cfg3_1024_tsc.json 2>> gpurun_out/bench_r1_uniform.err
python bench.py --steps 5 --warmup 3 --workload 256_pcs --no-cpu-baseline > gpurun_out/bench_r1_256_pcs.json 2>> gpurun_out/bench_r1_uniform.err
python profiles/ring_variants.py 512 > gpurun_out/ring_variants_512.txt 2>&1
python profiles/ring_variants.py 1024 > gpurun_out/ring_variants_1024.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
tail -c 600 gpurun_out/bench_r1_uniform.json; tail -c 300 gpurun_out/bench_r1_reference.json
