#!/bin/bash
# Round 2, first GPU check of the engine refactor: GPU tests, smoke, default bench, drop-in binary,
# RED-rate patterns.
set -u
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt
lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/${TAG}_gpu.txt
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.txt 2>&1; echo "smoke exit $?" >> $O/smoke_$TAG.txt
timeout 600 python bench.py > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err
timeout 120 python - > $O/red_rate_$TAG.txt 2>&1 <<'PY'
import ctypes as C
from neutral_b200.host import load_library
lib = load_library()
r = C.c_double()
print("pattern footprint_MiB reductions_per_s")
for p in (3, 1, 0, 5, 6):
    for mib in (16, 64, 128, 256):
        lib.nb200_microbench_red(p, mib << 20, 1000, C.byref(r))
        print(p, mib, "%.3e" % r.value)
PY
( cd build/run/neutral && NB200_RESULTS_JSON=$PWD/../../../$O/dropin_${TAG}_csp.json timeout 120 ./neutral.b200 problems/csp.params ) > $O/dropin_${TAG}_csp.txt 2>&1
tail -15 $O/pytest_gpu_$TAG.txt; tail -2 $O/smoke_$TAG.txt; tail -4 $O/dropin_${TAG}_csp.txt; cut -c1-600 $O/bench_${TAG}_csp.json; tail -3 $O/bench_${TAG}_csp.err; cat $O/red_rate_$TAG.txt
