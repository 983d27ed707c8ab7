#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "sharded_over_gpus and mixed_small" > $O/sanitizer_multi_n2.txt 2>&1; echo "sanitizer exit $?" >> $O/sanitizer_multi_n2.txt
grep "^========= Program hit\|^========= Invalid\|^========= Error\|ERROR SUMMARY\|passed\|failed\|sanitizer exit" $O/sanitizer_multi_n2.txt | sort | uniq -c | head
