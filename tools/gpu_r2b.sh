#!/bin/bash
# Round 2: bench.py with progress markers and a watchdog, then the warp timeline of csp / split.
set -u
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
NB200_BENCH_VERBOSE=1 NB200_BENCH_WATCHDOG=240 timeout 300 python bench.py > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err
echo "bench exit $?"; tail -40 $O/bench_${TAG}_csp.err
for s in 2 5 8; do
  NB200_LIB=libneutral_b200.trace.so timeout 120 python tools/warp_trace.py csp --step $s > $O/warp_trace_${TAG}_csp_step$s.txt 2>&1
done
NB200_LIB=libneutral_b200.trace.so timeout 120 python tools/warp_trace.py split --step 1 > $O/warp_trace_${TAG}_split.txt 2>&1
NB200_LIB=libneutral_b200.trace.so timeout 120 python tools/warp_trace.py scatter --step 1 --bins 12 > $O/warp_trace_${TAG}_scatter.txt 2>&1
cat $O/warp_trace_${TAG}_csp_step5.txt; head -8 $O/warp_trace_${TAG}_split.txt
cut -c1-300 $O/bench_${TAG}_csp.json
