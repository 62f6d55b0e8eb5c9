set -x
mkdir -p gpurun_out
./profiles/microbench/atoms_pattern > gpurun_out/r2c_atoms_pattern.txt 2>&1
cat gpurun_out/r2c_atoms_pattern.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_fixed -s 1 -c 1 -o gpurun_out/r2c_fixed_pcs python profiles/ncu_deposit.py 512 PCS 2 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deposit_tile -s 1 -c 1 -o gpurun_out/r2c_float_pcs python profiles/ncu_deposit.py 512 PCS 1 >> gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
