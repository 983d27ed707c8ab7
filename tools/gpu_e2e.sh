#!/bin/bash
# e2e experiments: tools/gpu_e2e.sh "<opts1>" "<opts2>" ...
set -u
O=gpurun_out; mkdir -p $O
i=0
for opts in "$@"; do
  timeout 300 python bench.py --deck ${DECK:-csp} --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --opts "$opts" \
      > $O/e2e_${i}.json 2> $O/e2e_${i}.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("$O/e2e_${i}.json") if l.startswith("{")][0]
    print("[$opts] resident %.4e (%.2f ms; hist share %.3f sort %.3f)  e2e %.4e (%.2f ms; hist %.2f sort %.2f)  clk %s %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_share_of_step"], d["roofline"]["sort_phase_share_of_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["history_kernel_ms_per_step"], d["e2e"]["sort_phase_ms_per_step"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("[$opts] failed", e); print(open("$O/e2e_${i}.err").read()[-1500:])
PY
  i=$((i+1))
done
