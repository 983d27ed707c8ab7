#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NB200_E2E_TRACE=1 timeout 300 $TR bench.py --gpus $N --steps 8 --warmup 3 --no-parity > $O/bench_e2e_trace_n$N.json 2> $O/bench_e2e_trace_n$N.err
grep "e2e host trace" $O/bench_e2e_trace_n$N.err
NB200_E2E_TRACE=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-parity --no-decks --no-cpu-baseline > $O/bench_e2e_trace_n1.json 2> $O/bench_e2e_trace_n1.err
grep "e2e host trace" $O/bench_e2e_trace_n1.err
python - <<PY
import json
for n in ($N, 1):
    j = [json.loads(l) for l in open("$O/bench_e2e_trace_n%d.json" % n) if l.startswith("{")][0]
    print(n, "value %.4e e2e %.4e ms/step %.2f e2e ms/step %.2f hist %.2f sort %.2f" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["ms_per_step"], j["e2e"]["history_kernel_ms_per_step"], j["e2e"]["sort_phase_ms_per_step"]))
PY
