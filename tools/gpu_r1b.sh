#!/bin/bash
# Session-4 A/B run on one B200: GPU parity tests (incl. tests/test_gpu_variants.py), then
# bench A/Bs: exact-one-half shortcut on/off, side-stream staging on/off, occupancy probes.
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_r1b.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_r1b.txt
tail -4 $O/pytest_gpu_r1b.txt
run() {  # tag deck opts [env]
  local tag=$1 deck=$2 opts=$3
  timeout 300 python bench.py --deck $deck --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --opts "$opts" \
      > $O/bench_${tag}_$deck.json 2> $O/bench_${tag}_$deck.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/bench_${tag}_$deck.json") if l.startswith("{")][0]
    r=d["roofline"]
    print("$tag $deck %.4e ev/s  ms/step %.2f  hist %.3f sort %.3f clk %s %s" % (d["value"], d["ms_per_step"], r["kernel_share_of_step"], r["sort_phase_share_of_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("$tag $deck failed", e); print(open("$O/bench_${tag}_$deck.err").read()[-1500:])
PY
}
for deck in csp split scatter stream; do run base $deck ""; done
for deck in csp split; do run nooverlap $deck "stage_overlap=0"; done
export NB200_LIB=libneutral_b200.nohalf.so
for deck in csp split scatter; do run nohalf $deck ""; done
unset NB200_LIB
# occupancy probes: 5 / 4 / 3 resident CTAs per SM (227 KB of shared memory per SM)
for pad in 38000 50000; do
  for deck in stream scatter; do run pad$pad $deck "history_smem_pad=$pad"; done
done
run pad68000 stream "history_smem_pad=68000"
run base2 csp ""
timeout 120 python tools/step_breakdown.py csp > $O/steps_r1b_csp.txt 2>&1; cat $O/steps_r1b_csp.txt
