#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for o in step_graph=1 step_graph=0; do
  timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-parity --opts $o > $O/bench_graph8_$o.json 2> $O/bench_graph8_$o.err
  python - "$O/bench_graph8_$o.json" "$o" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
    e = j["e2e"]
    print("bench x8", sys.argv[2], "value %.4e ms/step %.3f e2e %.4e (%.2f ms) e2e sort %.2f hist %.2f" % (j["value"], j["ms_per_step"], e["value"], e["ms_per_step"], e["sort_phase_ms_per_step"], e["history_kernel_ms_per_step"]))
except Exception as ex:
    print("bench x8", sys.argv[2], "failed", ex)
PY
done
