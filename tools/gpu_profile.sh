#!/bin/bash
# ncu evidence for the history kernel: launch list of a bench run + one full capture.
# usage: tools/gpu_profile.sh <tag> [deck] [skip] [count] [opts]
set -u
TAG=${1:-v1}; DECK=${2:-csp}; SKIP=${3:-13}; COUNT=${4:-2}; OPTS=${5:-}
mkdir -p gpurun_out
BENCH="python bench.py --deck $DECK --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --opts=$OPTS"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_${TAG}_${DECK}.csv $BENCH > gpurun_out/ncu_bench_${TAG}_${DECK}.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_history -s $SKIP -c $COUNT \
  -f -o gpurun_out/prof_${TAG}_${DECK} $BENCH >> gpurun_out/ncu_bench_${TAG}_${DECK}.log 2>&1
ls -la gpurun_out | grep prof_${TAG}
