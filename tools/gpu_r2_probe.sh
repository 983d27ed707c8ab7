#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
for lib in libneutral_b200.so libneutral_b200.probe1.so libneutral_b200.probe2.so; do
  for d in stream csp split; do
    echo "== $lib $d: $(NB200_LIB=$lib timeout 120 python tools/step_breakdown.py $d --repeat 3 2>&1 | tail -1)"
  done
done | tee $O/probe_tally.txt
NB200_LIB=libneutral_b200.probe1.so timeout 120 python tools/step_breakdown.py csp --repeat 3 > $O/probe_tally_csp_steps_noatomic.txt 2>&1
cat $O/probe_tally_csp_steps_noatomic.txt
