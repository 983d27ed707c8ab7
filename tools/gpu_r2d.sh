#!/bin/bash
# Round 2: regression (GPU tests), the default bench line, and the dispatch-interleave sweep.
set -u
TAG=${1:-r2d}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_$TAG.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.txt
tail -4 $O/pytest_gpu_$TAG.txt
NB200_BENCH_VERBOSE=1 NB200_BENCH_WATCHDOG=400 timeout 500 python bench.py > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err
echo "bench exit $?"; tail -12 $O/bench_${TAG}_csp.err
for q in 0 1 2 3 4 8; do
  for d in split csp; do
    timeout 120 python tools/step_breakdown.py $d --opts interleave=$q --repeat 3 > $O/steps_${TAG}_${d}_q$q.txt 2>&1
    echo "$d interleave=$q: $(tail -1 $O/steps_${TAG}_${d}_q$q.txt)"
  done
done
timeout 200 python bench.py --deck split_scaled --particles 12500000 --no-e2e --no-cpu-baseline --no-decks --steps 3 --warmup 2 > $O/bench_${TAG}_split125.json 2> $O/bench_${TAG}_split125.err
timeout 200 python bench.py --deck split_scaled --particles 12500000 --no-e2e --no-cpu-baseline --no-decks --steps 3 --warmup 2 --opts interleave=2 > $O/bench_${TAG}_split125_q2.json 2> $O/bench_${TAG}_split125_q2.err
python - <<PY
import json
for n in ["csp", "split125", "split125_q2"]:
    try:
        j=[json.loads(l) for l in open("$O/bench_${TAG}_%s.json" % n) if l.startswith("{")][0]
        print(n, "value %.4e" % j["value"], "ms/step %.3f" % j["ms_per_step"], "e2e", j.get("e2e", {}).get("value"), "parity", j.get("parity", {}).get("ok"), j.get("parity", {}).get("fixture"))
        if n == "csp":
            print(" roofline", {k: j["roofline"][k] for k in ("achieved", "peak", "frac", "avg_launch_ms", "kernel_share_of_step", "sort_phase_share_of_step")})
            print(" decks", {k: (v.get("value"), v.get("counts_match_reference")) for k, v in j.get("decks", {}).items()})
            print(" cpu", j.get("cpu_baseline"))
    except Exception as e:
        print(n, "failed", e)
PY
