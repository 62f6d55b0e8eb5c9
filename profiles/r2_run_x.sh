set -x
mkdir -p gpurun_out
cp pylians_b200/lib/libpylians_b200.so /tmp/default.so
for v in default d2c2 d4c2; do
  if [ $v = default ]; then cp /tmp/default.so pylians_b200/lib/libpylians_b200.so; else cp pylians_b200/lib/variants/$v.so pylians_b200/lib/libpylians_b200.so; fi
  echo "== variant $v" | tee -a gpurun_out/r2x_deposit_ab.txt
  timeout 400 python profiles/deposit_ab.py 512 1024 2>&1 | grep "TSC\|PCS" | grep "kernel=2" | tee -a gpurun_out/r2x_deposit_ab.txt
done
cp /tmp/default.so pylians_b200/lib/libpylians_b200.so
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "ma_vs_oracle_both or clustered or ma_weights or zeldovich_lattice" 2>&1 | tail -3
