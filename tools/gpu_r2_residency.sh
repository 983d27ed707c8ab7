#!/bin/bash
# Facet throughput of the event loop against resident warps per SM: csp's collision-free
# timesteps 1-2 with banks that fill 1, 2, 3, 4, 5, 6 CTA slots of every SM exactly once (one wave).
set -u
O=gpurun_out; mkdir -p $O
{
for ctas in 1 2 3 4 5 6 12; do
  n=$((148 * 128 * ctas))
  echo "== $ctas CTAs per SM ($n particles)"
  timeout 100 python tools/step_breakdown.py csp --particles $n --repeat 3 2>&1 | sed -n 3,5p
done
} > $O/residency_curve.txt 2>&1
cat $O/residency_curve.txt
