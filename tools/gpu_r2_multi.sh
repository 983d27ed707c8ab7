#!/bin/bash
# Round 2, multi-GPU check on N GPUs of one box: the sharded-bank tests (one process, N GPUs),
# the drop-in binary under NB200_NGPUS, and bench.py under torchrun (weak and strong scaling,
# csp and split, both collective flavours), each line carrying its parity object.
# usage: tools/gpu_r2_multi.sh <tag> <ngpus> [quick]
set -u
TAG=${1:-r2m}
N=${2:-2}
QUICK=${3:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt
nvidia-smi topo -m >> $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -k "sharded or several or pipelined" > $O/pytest_${TAG}.txt 2>&1; echo "pytest exit $?" >> $O/pytest_${TAG}.txt
tail -6 $O/pytest_${TAG}.txt
( cd build/run/neutral && NB200_NGPUS=$N timeout 300 ./neutral.b200 problems/split.params ) > $O/dropin_${TAG}_split_n$N.txt 2>&1
( cd build/run/neutral && NB200_NGPUS=$N timeout 300 ./neutral.b200 problems/csp.params ) > $O/dropin_${TAG}_csp_n$N.txt 2>&1
grep -E "Step time|Facets|PASSED|FAILED|sharded|Final" $O/dropin_${TAG}_csp_n$N.txt | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() {  # name, bench args...
  local name=$1; shift
  timeout 600 $TR bench.py --gpus $N "$@" > $O/bench_${TAG}_$name.json 2> $O/bench_${TAG}_$name.err
  python - "$O/bench_${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    j = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
    p = j.get("parity", {})
    print(sys.argv[2], "value %.4e" % j["value"], "e2e %.4e" % j.get("e2e", {}).get("value", 0), "ms/step %.2f" % j["ms_per_step"],
          "hist ms %.3f (slowest %.3f)" % (j["roofline"]["avg_launch_ms"], j["roofline"]["slowest_rank_avg_launch_ms"]),
          "parity", p.get("ok"), p.get("counts_match"), p.get("bank_bit_identical"), p.get("tally_block_max_rel_err"), j["run"]["parallelism"][:60])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run csp_weak --steps 5 --warmup 3
run csp_strong --steps 5 --warmup 3 --scaling strong
run split_weak --steps 5 --warmup 3 --deck split
if [ -z "$QUICK" ]; then
  run split_strong --steps 5 --warmup 3 --deck split --scaling strong
  run csp_weak_nccl --steps 5 --warmup 3 --opts collective=0
  run csp_weak_noe2e_20 --steps 20 --warmup 5 --no-e2e
fi
tail -3 $O/bench_${TAG}_*.err | tail -20
