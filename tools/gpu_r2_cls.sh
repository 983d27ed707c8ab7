#!/bin/bash
# Round 2, session 2: the sort's class rule (collides_soon) A/B on one box, the full GPU test
# suite with step_graph=1 as the default, and a default bench line.
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_s2.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_s2.txt
tail -4 $O/pytest_gpu_s2.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for rep in 1 2; do
  for lib in libneutral_b200.cls0.so libneutral_b200.so; do
    for d in split csp scatter; do
      echo "== rep $rep $lib $d: $(NB200_LIB=$lib timeout 200 python tools/step_breakdown.py $d --repeat 3 2>&1 | tail -1)"
    done
  done
done | tee $O/cls_ab.txt
NB200_LIB=libneutral_b200.cls0.so timeout 100 python tools/step_breakdown.py split --repeat 2 > $O/steps_s2_split_cls0.txt 2>&1
timeout 100 python tools/step_breakdown.py split --repeat 2 > $O/steps_s2_split.txt 2>&1
timeout 100 python tools/step_breakdown.py csp --repeat 2 > $O/steps_s2_csp.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s2.json 2> $O/bench_s2.err; echo "bench exit $?"
python - <<'PY'
import json
j=[json.loads(l) for l in open("gpurun_out/bench_s2.json") if l.startswith("{")][0]
e=j["e2e"]; p=j.get("parity",{})
print("bench value %.4e ms %.3f e2e %.4e (%.2f ms) roofline frac %.3f parity %s decks %s" % (j["value"], j["ms_per_step"], e["value"], e["ms_per_step"], j["roofline"]["frac"], {k:v for k,v in p.items() if isinstance(v,bool)}, {k:(v.get("value") if isinstance(v,dict) else v) for k,v in j.get("decks",{}).items()}))
PY
