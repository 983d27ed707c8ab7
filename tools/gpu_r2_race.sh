#!/bin/bash
# racecheck (shared-memory hazards) over one small deck in the default configuration and with
# the staggered dispatch map.
set -u
O=gpurun_out; mkdir -p $O
timeout 50 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "test_device_flavour_matches_oracle_every_step and mixed_small and (pipeline-stagger or pipeline-calls)" > $O/racecheck_r02d.txt 2>&1; echo "racecheck exit $?" >> $O/racecheck_r02d.txt
tail -5 $O/racecheck_r02d.txt
