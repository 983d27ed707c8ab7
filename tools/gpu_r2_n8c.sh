#!/bin/bash
set -u
TAG=${1:-r2n8c}
N=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for steps in "10 3" "20 5"; do set -- $steps
timeout 400 $TR bench.py --gpus $N --steps $1 --warmup $2 > $O/bench_${TAG}_csp_weak_s$1.json 2> $O/bench_${TAG}_csp_weak_s$1.err
python - "$O/bench_${TAG}_csp_weak_s$1.json" <<'PY'
import json, sys
j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
p, e = j.get("parity", {}), j.get("e2e", {})
print("csp weak x8 steps %d value %.4e e2e %.4e ms/step %.2f e2e ms/step %.2f hist %.3f e2e sort %.2f parity %s" % (j["steps"], j["value"], e["value"], j["ms_per_step"], e["ms_per_step"], j["roofline"]["avg_launch_ms"], e["sort_phase_ms_per_step"], p.get("ok")))
PY
done
