#!/bin/bash
# First-contact GPU session: environment probes, parity tests, smoke, bench, launch list.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
{
  echo "== probes"; ls -d /root/reference 2>&1; nproc; lscpu | grep -E 'Model name|Socket|Thread|Core' ; free -g | head -2
  nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv
  nvidia-smi topo -m 2>/dev/null | head -12
  ls oracle/_ref 2>&1
} > gpurun_out/probes.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.txt 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -3; cat gpurun_out/bench.txt; tail -5 gpurun_out/bench.err
