#!/bin/bash
# Round 2, final 8-GPU lines of the build with step_graph=1 and the sort's class rule as
# defaults: csp weak (the driver's scaling line), split weak, split scaled to 1e8 - each with
# its parity object against the reference fixture of its own particle count.
set -u
TAG=${1:-r2n8c}
N=${2:-8}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() {  # name, bench args...
  local name=$1; shift
  timeout 400 $TR bench.py --gpus $N "$@" > $O/bench_${TAG}_$name.json 2> $O/bench_${TAG}_$name.err
  python - "$O/bench_${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
    p = j.get("parity", {})
    e = j.get("e2e", {})
    print(sys.argv[2], "value %.4e" % j["value"], "e2e %.4e" % e.get("value", 0), "ms/step %.2f" % j["ms_per_step"],
          "e2e ms %.2f" % e.get("ms_per_step", 0),
          "hist ms %.3f (slowest %.3f)" % (j["roofline"]["avg_launch_ms"], j["roofline"]["slowest_rank_avg_launch_ms"]),
          "e2e sort ms %.2f hist %.2f" % (e.get("sort_phase_ms_per_step", 0), e.get("history_kernel_ms_per_step", 0)),
          "parity", p.get("ok"), p.get("counts_match"), p.get("bank_bit_identical"), p.get("tally_block_max_rel_err"), p.get("fixture", "")[:40])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run csp_weak --steps 20 --warmup 5
run split_weak --steps 10 --warmup 3 --deck split
run split_scaled --steps 3 --warmup 2 --deck split_scaled --particles 12500000 --no-e2e
tail -2 $O/bench_${TAG}_*.err | grep -v "^$" | tail -8
