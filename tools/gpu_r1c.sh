#!/bin/bash
# A/B run: parity tests, then the four decks (resident numbers only).
set -u
O=gpurun_out
TAG=${1:-r1c}
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$TAG.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.txt
tail -15 $O/pytest_gpu_$TAG.txt
run() {  # tag deck opts
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
for deck in csp split scatter stream; do run $TAG $deck ""; done
for lib in ${NB200_VARIANTS:-}; do
  export NB200_LIB=libneutral_b200.$lib.so
  for deck in csp split scatter stream; do run ${TAG}_$lib $deck ""; done
  unset NB200_LIB
done
timeout 120 python tools/step_breakdown.py csp > $O/steps_${TAG}_csp.txt 2>&1; cat $O/steps_${TAG}_csp.txt
