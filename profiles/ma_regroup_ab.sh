# A/B: tile kernel with particles regrouped by shared-memory bank (PYLB_MA_REGROUP=1, default) vs arrival order
for wl in cfg2_512_cic 256_pcs; do for f in 0 1; do echo "workload=$wl PYLB_MA_REGROUP=$f"; PYLB_MA_REGROUP=$f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload $wl | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ms/step %.3f  deposit %.3f ms  tile kernel %.3f ms  pk %.3f ms' % (d['ms_per_step'], d['stages']['deposit_ms'], d['kernels']['tile_ms'], d['stages']['pk_ms']))"; done; done
echo "fixed point + regroup"; PYLB_MA_FIXED=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ms/step %.3f  deposit %.3f ms  tile kernel %.3f ms' % (d['ms_per_step'], d['stages']['deposit_ms'], d['kernels']['tile_ms']))"
