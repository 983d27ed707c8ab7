#!/bin/bash
# Final check of a build on one B200: the whole GPU suite, smoke, the default bench line with the
# driver's arguments.   usage: tools/gpu_r2_verify.sh <tag>
set -u
TAG=${1:-verify}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.txt
tail -3 $O/pytest_gpu_$TAG.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.txt 2>&1; echo "smoke exit $?" >> $O/smoke_$TAG.txt; tail -2 $O/smoke_$TAG.txt
timeout 400 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err; echo "bench exit $?"
python - <<PY
import json
j=[json.loads(l) for l in open("$O/bench_${TAG}_csp.json") if l.startswith("{")][0]
e=j["e2e"]; p=j.get("parity",{})
print("value %.4e ms %.3f e2e %.4e (%.2f ms) frac %.3f parity %s decks %s cpu %s" % (j["value"], j["ms_per_step"], e["value"], e["ms_per_step"], j["roofline"]["frac"], p.get("ok"), {k:"%.3e"%v["value"] for k,v in j.get("decks",{}).items()}, j.get("cpu_baseline",{}).get("value")))
PY
