# span schedule sweep of ring2: PYLB_RING2_SHARE = share of the rows in the first (long) span of every CTA
for N in 512 1024; do for f in 1.0 0.8 0.6 0.4; do echo "N=$N share=$f"; PYLB_RING2_SHARE=$f timeout 120 python profiles/ring_variants.py $N 2>&1 | head -3; done; done
PYLB_RING2_SHARE=1.0 PYLB_RING2_TRACE=gpurun_out/ring2_trace_static.txt timeout 120 python profiles/ring_profile.py 512 2
PYLB_RING2_SHARE=0.6 PYLB_RING2_TRACE=gpurun_out/ring2_trace_guided.txt timeout 120 python profiles/ring_profile.py 512 2
