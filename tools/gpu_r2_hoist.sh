#!/bin/bash
# A/B: the absorbed lanes' path-sample division taken before the absorb / scatter branch
# (-DNB_HOIST_ABSORB_MFP) against the product library, same box.
set -u
O=gpurun_out; mkdir -p $O
NB200_LIB=libneutral_b200.hoist.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "pipeline and not ieee" 2>&1 | tail -1
{
for rep in 1 2; do
  for lib in libneutral_b200.so libneutral_b200.hoist.so; do
    for d in split csp scatter; do
      echo "== rep $rep $lib $d: $(NB200_LIB=$lib timeout 200 python tools/step_breakdown.py $d --repeat 3 2>&1 | tail -1)"
    done
  done
done
} > $O/hoist_ab.txt 2>&1
cat $O/hoist_ab.txt
