#!/bin/bash
# Round-2 evidence on one B200: GPU tests, smoke, the bench lines (default csp with every
# sub-object, the reference arm), per-timestep breakdowns, the ncu launch list of a bench run,
# full ncu captures of the history kernel summarised ON the box (text comes back, the reports
# stay there), a compute-sanitizer pass, and the unmodified reference driver linked against the
# library. Everything lands in gpurun_out/.   usage: tools/gpu_r2_final1.sh <tag>
set -u
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt
lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/${TAG}_gpu.txt
timeout 900 python -m pytest tests -q -m gpu > $O/pytest_gpu_$TAG.txt 2>&1; echo "pytest exit $?" >> $O/pytest_gpu_$TAG.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.txt 2>&1; echo "smoke exit $?" >> $O/smoke_$TAG.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
for d in csp split; do timeout 120 python tools/step_breakdown.py $d > $O/steps_${TAG}_$d.txt 2>&1; done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-decks --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_${TAG}_csp.csv $B > $O/ncu_${TAG}.log 2>&1
K=_ZN2nb9k_historyILb1ELb0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -s 13 -c 2 \
  -f -o /tmp/prof_csp $B >> $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -c 1 \
  -f -o /tmp/prof_stream $B --deck stream >> $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -c 1 \
  -f -o /tmp/prof_split $B --deck split >> $O/ncu_${TAG}.log 2>&1
for d in csp stream split; do
  python tools/ncu_summary.py /tmp/prof_$d.ncu-rep $O/ncu_${TAG}_$d.txt > /dev/null 2>&1
done
python tools/ncu_lines.py /tmp/prof_csp.ncu-rep $K > $O/ncu_${TAG}_csp_lines.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine.py tests/test_gpu_variants.py -x -q -k "(pipeline and not ieee) or pipelined or tables_edited or headroom or host_mirror or per_bank" \
  > $O/sanitizer_$TAG.txt 2>&1; echo "sanitizer exit $?" >> $O/sanitizer_$TAG.txt
( cd build/run/neutral && timeout 120 ./neutral.b200 problems/csp.params ) > $O/dropin_${TAG}_csp.txt 2>&1
python - <<PY
import json
for d in ["csp","reference"]:
    try:
        j=[json.loads(l) for l in open("$O/bench_${TAG}_%s.json"%d) if l.startswith("{")][0]
        print(d, "%.4e"%j["value"], "e2e %.4e"%j["e2e"]["value"], "ms/step %.2f"%j["ms_per_step"], j.get("clocks"), j.get("parity",{}).get("ok"), j.get("cpu_baseline",{}).get("sample"))
    except Exception as e:
        print(d, "failed", e)
PY
tail -3 $O/pytest_gpu_$TAG.txt; tail -2 $O/smoke_$TAG.txt; tail -4 $O/sanitizer_$TAG.txt; grep -E "Step time|PASSED" $O/dropin_${TAG}_csp.txt | head -4
