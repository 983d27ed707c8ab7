#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "graph" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_engine.py -q -m gpu -x -k "pipelined" 2>&1 | tail -2
for rep in 1 2; do
for o in step_graph=0 step_graph=1; do
  for d in csp split; do
    echo "== rep $rep $o $d: $(timeout 120 python tools/step_breakdown.py $d --opts $o --repeat 3 2>&1 | tail -1)"
  done
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-decks --no-parity --opts $o > $O/bench_graph_$o.json 2> $O/bench_graph_$o.err
  python - "$O/bench_graph_$o.json" "$o" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
    e = j["e2e"]
    print("   bench", sys.argv[2], "value %.4e ms/step %.3f e2e %.4e (%.2f ms) sort share %.4f launches %d" % (j["value"], j["ms_per_step"], e["value"], e["ms_per_step"], j["roofline"]["sort_phase_share_of_step"], j["gpu_launches"]))
except Exception as ex:
    print("   bench", sys.argv[2], "failed", ex)
PY
done
done
tail -3 $O/bench_graph_step_graph=1.err
