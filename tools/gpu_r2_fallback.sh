#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
NB200_TEST_IPC_FAIL=1 timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > $O/bench_fallback_n2.json 2> $O/bench_fallback_n2.err
python - <<'PY'
import json
j = [json.loads(l) for l in open("gpurun_out/bench_fallback_n2.json") if l.startswith("{")][0]
print("value %.4e ms/step %.2f parity %s | %s" % (j["value"], j["ms_per_step"], j["parity"]["ok"], j["run"]["parallelism"][:160]))
PY
tail -3 $O/bench_fallback_n2.err
