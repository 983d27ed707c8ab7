#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -X faulthandler -m pytest tests/test_gpu_engine.py -x -s -q -m gpu -k "sharded or several" > $O/pytest_dbg4.txt 2>&1; echo "exit $?" >> $O/pytest_dbg4.txt
grep -v "^Particles\|^$\|Step time\|Wallclock\|Facets\|Collisions\|Events\|Iteration" $O/pytest_dbg4.txt | tail -30 | cut -c1-250
( cd build/run/neutral && NB200_NGPUS=4 timeout 300 ./neutral.b200 problems/split.params ) 2>&1 | grep -E "Step time|Allocated|Final g"
( cd build/run/neutral && NB200_NGPUS=1 timeout 300 ./neutral.b200 problems/csp.params ) 2>&1 | grep -E "Step time" | head -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_dbg4_csp_weak.json 2> $O/bench_dbg4_csp_weak.err
python - <<'PY'
import json
j = [json.loads(l) for l in open("gpurun_out/bench_dbg4_csp_weak.json") if l.startswith("{")][0]
print("csp weak x4 value %.4e e2e %.4e ms/step %.2f parity %s" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["parity"]["ok"]))
PY
