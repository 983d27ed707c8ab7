#!/bin/bash
# Same-box A/B of library variants: tools/gpu_ab.sh "<lib suffixes, '-' = product>" "<decks>" [rounds]
set -u
O=gpurun_out; mkdir -p $O
LIBS=${1:--}; DECKS=${2:-csp split scatter}; ROUNDS=${3:-2}
for r in $(seq 1 $ROUNDS); do
 for deck in $DECKS; do
  for lib in $LIBS; do
    if [ "$lib" = "-" ]; then unset NB200_LIB; else export NB200_LIB=libneutral_b200.$lib.so; fi
    timeout 300 python bench.py --deck $deck --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline --no-e2e \
      > $O/ab_${lib}_${deck}_$r.json 2> $O/ab_${lib}_${deck}_$r.err
    python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/ab_${lib}_${deck}_$r.json") if l.startswith("{")][0]
    print("round $r  %-8s %-8s %.4e ev/s  ms/step %.3f  clk %s %s" % ("$lib", "$deck", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("$lib $deck failed", e); print(open("$O/ab_${lib}_${deck}_$r.err").read()[-800:])
PY
  done
 done
done
