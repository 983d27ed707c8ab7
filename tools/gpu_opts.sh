#!/bin/bash
# usage: tools/gpu_opts.sh <tag> "<opts1>" "<opts2>" ... ; env DECKS="csp stream" STEPS=5
set -u
O=gpurun_out; mkdir -p $O
TAG=$1; shift
i=0
for opts in "$@"; do
  for deck in ${DECKS:-csp stream}; do
    timeout 300 python bench.py --deck $deck --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-e2e --opts "$opts" \
      > $O/bench_${TAG}_${i}_$deck.json 2> $O/bench_${TAG}_${i}_$deck.err
    python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/bench_${TAG}_${i}_$deck.json") if l.startswith("{")][0]
    r=d["roofline"]
    print("$TAG [$opts] $deck %.4e ev/s  ms/step %.2f  hist %.3f sort %.3f clk %s %s" % (d["value"], d["ms_per_step"], r["kernel_share_of_step"], r["sort_phase_share_of_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("$TAG [$opts] $deck failed", e); print(open("$O/bench_${TAG}_${i}_$deck.err").read()[-1500:])
PY
  done
  i=$((i+1))
done
