#!/bin/bash
# Round 2 on 8 GPUs: sharded-bank tests (one process, 8 GPUs), the cost breakdown of the tally
# group, and the bench lines (weak / strong csp, split, split scaled to 1e8) with their parity.
set -u
TAG=${1:-r2n8}
N=${2:-8}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "sharded or several" > $O/pytest_${TAG}.txt 2>&1; tail -3 $O/pytest_${TAG}.txt
timeout 300 $TR tools/group_breakdown.py csp > $O/group_breakdown_${TAG}.txt 2>&1; grep "csp x" $O/group_breakdown_${TAG}.txt
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
          "hist ms %.3f (slowest %.3f)" % (j["roofline"]["avg_launch_ms"], j["roofline"]["slowest_rank_avg_launch_ms"]),
          "e2e sort ms %.2f" % e.get("sort_phase_ms_per_step", 0),
          "parity", p.get("ok"), p.get("counts_match"), p.get("bank_bit_identical"), p.get("tally_block_max_rel_err"), p.get("fixture", "")[:40])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run csp_weak --steps 5 --warmup 3
run csp_strong --steps 5 --warmup 3 --scaling strong
run split_scaled --steps 3 --warmup 2 --deck split_scaled --particles 12500000 --no-e2e
run split_weak --steps 5 --warmup 3 --deck split
run split_strong --steps 5 --warmup 3 --deck split --scaling strong --no-e2e
run csp_weak_every1 --steps 5 --warmup 3 --opts tally_reduce_every=1 --no-e2e
( cd build/run/neutral && NB200_NGPUS=$N timeout 300 ./neutral.b200 problems/split_scaled.params ) > $O/dropin_${TAG}_split_scaled.txt 2>&1
grep -E "Step time|Facets|Collisions|Final|validate|sharded|Allocated" $O/dropin_${TAG}_split_scaled.txt | tail -8
tail -2 $O/bench_${TAG}_*.err | grep -v "^$" | tail -12
