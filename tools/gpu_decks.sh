#!/bin/bash
# events/s of every BASELINE deck on one GPU (device-resident value + e2e), no CPU baseline.
set -u
mkdir -p gpurun_out
for deck in stream csp split scatter; do
  timeout 900 python bench.py --deck $deck --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${1:-v1}_$deck.json 2> gpurun_out/bench_${1:-v1}_$deck.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${1:-v1}_$deck.json"))
    print("$deck", "%.3e ev/s"%d["value"], "e2e %.3e"%d["e2e"]["value"], "ms/step %.1f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$deck failed", e)
PY
done
