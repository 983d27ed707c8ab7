#!/bin/bash
# What the collective costs the transport it runs beside (N GPUs): breakdown + knobs.
set -u
TAG=${1:-r2mb}
N=${2:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "sharded or several" 2>&1 | tail -2
timeout 300 $TR tools/group_breakdown.py csp > $O/group_breakdown_${TAG}_n$N.txt 2>&1; cat $O/group_breakdown_${TAG}_n$N.txt | grep "csp x" | tail -12
run() {  # name, env, bench args...
  local name=$1; local envs=$2; shift; shift
  env $envs timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 "$@" > $O/bench_${TAG}_$name.json 2> $O/bench_${TAG}_$name.err
  python - "$O/bench_${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
    p = j.get("parity", {})
    print(sys.argv[2], "value %.4e" % j["value"], "e2e %.4e" % j.get("e2e", {}).get("value", 0), "ms/step %.2f" % j["ms_per_step"],
          "hist ms %.3f (slowest %.3f)" % (j["roofline"]["avg_launch_ms"], j["roofline"]["slowest_rank_avg_launch_ms"]),
          "parity", p.get("ok"), p.get("tally_block_max_rel_err"))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run weak A=1
run weak_every1 A=1 --opts tally_reduce_every=1 --no-e2e
run strong A=1 --scaling strong
run split_weak A=1 --deck split
