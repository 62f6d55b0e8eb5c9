mkdir -p gpurun_out
cp pylians_b200/lib/libpylians_b200.so /tmp/default.so
for v in default pt0_256; do
  if [ $v = default ]; then cp /tmp/default.so pylians_b200/lib/libpylians_b200.so; else cp pylians_b200/lib/variants/$v.so pylians_b200/lib/libpylians_b200.so; fi
  echo "== variant $v" | tee -a gpurun_out/r2an_deposit_ab.txt
  timeout 400 python profiles/deposit_ab.py 1024 2>&1 | grep "NGP\|CIC kernel\|PCS kernel=2" | tee -a gpurun_out/r2an_deposit_ab.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short -k "deep_sort or ma_vs_oracle_both" 2>&1 | tail -2
cp /tmp/default.so pylians_b200/lib/libpylians_b200.so
