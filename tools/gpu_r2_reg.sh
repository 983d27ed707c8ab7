#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_reg.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_reg.txt
tail -4 $O/pytest_gpu_reg.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --no-cpu-baseline > $O/bench_reg.json 2> $O/bench_reg.err; echo "bench exit $?"; cut -c1-200 $O/bench_reg.json
