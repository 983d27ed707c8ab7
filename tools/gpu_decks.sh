#!/bin/bash
# events/s of every BASELINE deck on one GPU (device-resident value + e2e), no CPU baseline.
# usage: tools/gpu_decks.sh <tag> [opts] [decks...]
set -u
TAG=${1:-v1}; OPTS=${2:-}; shift 2 2>/dev/null
DECKS=${@:-stream csp split scatter}
mkdir -p gpurun_out
for deck in $DECKS; do
  timeout 900 python bench.py --deck $deck --steps 3 --warmup 3 --no-cpu-baseline --opts "$OPTS" > gpurun_out/bench_${TAG}_$deck.json 2> gpurun_out/bench_${TAG}_$deck.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$deck.json"))
    r=d["roofline"]
    print("$TAG $deck", "%.3e ev/s"%d["value"], "e2e %.3e"%d["e2e"]["value"], "ms/step %.1f"%d["ms_per_step"], "frac %.3f"%r["frac"], "hist %.3f sort %.3f"%(r["kernel_share_of_step"], r.get("sort_phase_share_of_step",0)), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$TAG $deck failed", e); print(open("gpurun_out/bench_${TAG}_$deck.err").read()[-2000:])
PY
done
