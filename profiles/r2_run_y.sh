mkdir -p gpurun_out
./profiles/microbench/atoms_pattern | tee gpurun_out/r2y_atoms_pattern.txt
timeout 400 python profiles/deposit_ab.py 512 2>&1 | grep "TSC\|PCS" | grep "kernel=2"
