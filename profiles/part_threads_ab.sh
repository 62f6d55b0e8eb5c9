# A/B: block-local counting sort passes with 512-thread CTAs (4096-particle chunks, 2 CTAs/SM) vs 256 (2048, 4 CTAs/SM)
for t in 512 256; do echo "PYLB_PART_THREADS=$t"; PYLB_PART_THREADS=$t python bench.py --steps 5 --warmup 3 --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ms/step %.3f  deposit %.3f ms  tile kernel %.3f ms  pk %.3f ms' % (d['ms_per_step'], d['stages']['deposit_ms'], d['kernels']['tile_ms'], d['stages']['pk_ms']))"; done
