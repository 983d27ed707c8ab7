#!/bin/bash
# Round 2 on 4 GPUs: the missing points of the split 1/2/4/8 table, csp strong scaling, and a
# compute-sanitizer memcheck pass over the sharded-bank tests (peer-memory kernels included).
set -u
TAG=${1:-r2n4}
N=4
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
    p, e = j.get("parity", {}), j.get("e2e", {})
    print(sys.argv[2], "value %.4e" % j["value"], "e2e %.4e" % e.get("value", 0), "ms/step %.2f" % j["ms_per_step"], "parity", p.get("ok"), p.get("fixture", "")[:44])
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run split_weak --steps 5 --warmup 3 --deck split
run split_strong --steps 5 --warmup 3 --deck split --scaling strong
run csp_strong --steps 5 --warmup 3 --scaling strong
run split_scaled_strong --steps 2 --warmup 1 --deck split_scaled --scaling strong --no-e2e
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "sharded_over_gpus and mixed_small" > $O/sanitizer_multi_${TAG}.txt 2>&1; echo "sanitizer exit $?" >> $O/sanitizer_multi_${TAG}.txt
tail -5 $O/sanitizer_multi_${TAG}.txt
