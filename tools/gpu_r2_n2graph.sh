#!/bin/bash
# Sharded banks in ONE process (NB200_NGPUS) with step_graph=1 as the default: the engine tests
# that skip on a one-GPU box, and the unmodified reference driver on 2 GPUs.
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x > $O/pytest_n2_graph.txt 2>&1; echo "pytest exit $?" >> $O/pytest_n2_graph.txt
tail -3 $O/pytest_n2_graph.txt
for d in split csp; do
( cd build/run/neutral && NB200_NGPUS=2 timeout 120 ./neutral.b200 problems/$d.params ) > $O/dropin_n2_graph_$d.txt 2>&1
grep -E "Facets|Collisions|PASSED|FAILED|sharded|Final" $O/dropin_n2_graph_$d.txt | tail -5
done
