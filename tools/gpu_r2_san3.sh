#!/bin/bash
# memcheck over the parity tests of the kernel configurations added in session 2 of round 2
# (staggered dispatch map, call-by-call submission beside the graph default).
set -u
O=gpurun_out; mkdir -p $O
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "stagger or calls" > $O/sanitizer_r02d.txt 2>&1; echo "sanitizer exit $?" >> $O/sanitizer_r02d.txt
tail -5 $O/sanitizer_r02d.txt
