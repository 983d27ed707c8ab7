#!/bin/bash
# Round 2, second pass on 8 GPUs: the sharded-bank tests with their output visible, the box's
# aggregate host<->device copy rate, the csp weak-scaling line, and the drop-in binary on 8 GPUs.
set -u
TAG=${1:-r2n8b}
N=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -X faulthandler -m pytest tests/test_gpu_engine.py -x -s -q -m gpu -k "sharded or several" > $O/pytest_${TAG}.txt 2>&1; echo "pytest exit $?" >> $O/pytest_${TAG}.txt
grep -v "^Particles\|^$\|Step time\|Wallclock\|Facets\|Collisions\|Events\|Iteration" $O/pytest_${TAG}.txt | tail -25 | cut -c1-250
timeout 200 $TR tools/pcie_probe_all.py > $O/pcie_probe_all_${TAG}.txt 2>&1; grep ranks $O/pcie_probe_all_${TAG}.txt
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${TAG}_csp_weak.json 2> $O/bench_${TAG}_csp_weak.err
python - "$O/bench_${TAG}_csp_weak.json" <<'PY'
import json, sys
j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
p, e = j.get("parity", {}), j.get("e2e", {})
print("csp weak x8 value %.4e e2e %.4e ms/step %.2f e2e ms/step %.2f hist %.3f parity %s" % (j["value"], e["value"], j["ms_per_step"], e["ms_per_step"], j["roofline"]["avg_launch_ms"], p.get("ok")))
PY
for d in csp split; do
  ( cd build/run/neutral && NB200_NGPUS=$N timeout 300 ./neutral.b200 problems/$d.params ) > $O/dropin_${TAG}_${d}_n$N.txt 2>&1
  grep -E "Step time|PASSED|FAILED|Final g" $O/dropin_${TAG}_${d}_n$N.txt | head -5
done
