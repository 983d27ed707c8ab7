#!/bin/bash
# Where do the delayed colliders of the stagger experiment run, and for how long?
set -u
O=gpurun_out; mkdir -p $O
for at in 0 40; do
  NB200_LIB=libneutral_b200.trace.so timeout 120 python tools/warp_trace.py csp --step 6 --opts stagger_at=$at,stagger_min=40 > $O/warp_trace_stagger${at}_csp_step6.txt 2>&1
  head -8 $O/warp_trace_stagger${at}_csp_step6.txt
done
