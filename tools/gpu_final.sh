#!/bin/bash
# Round-end evidence on one B200: bench lines of every deck, the reference arm, per-timestep
# breakdowns, the ncu launch list of a bench run and full captures of the history kernel, and
# the unmodified reference driver linked against the library. Everything lands in gpurun_out/.
# usage: tools/gpu_final.sh <tag>
set -u
TAG=${1:-final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt
lscpu | grep -E 'Model name|^CPU\(s\)' >> $O/${TAG}_gpu.txt
timeout 600 python bench.py > $O/bench_${TAG}_csp.json 2> $O/bench_${TAG}_csp.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err
for d in stream split scatter; do
  timeout 300 python bench.py --deck $d --no-cpu-baseline > $O/bench_${TAG}_$d.json 2> $O/bench_${TAG}_$d.err
done
for d in csp split; do timeout 120 python tools/step_breakdown.py $d > $O/steps_${TAG}_$d.txt 2>&1; done
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_${TAG}_csp.csv $B > $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -s 13 -c 2 \
  -f -o $O/prof_${TAG}_csp $B >> $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -c 1 \
  -f -o $O/prof_${TAG}_stream $B --deck stream >> $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -c 1 \
  -f -o $O/prof_${TAG}_split $B --deck split >> $O/ncu_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_history -c 1 \
  -f -o $O/prof_${TAG}_scatter python tools/step_breakdown.py scatter --particles 2000000 --repeat 1 >> $O/ncu_${TAG}.log 2>&1
( cd build/run/neutral && timeout 120 ./neutral.b200 problems/csp.params ) > $O/dropin_${TAG}_csp.txt 2>&1
python - <<PY
import json
for d in ["csp","stream","split","scatter","reference"]:
    try:
        j=[json.loads(l) for l in open("$O/bench_${TAG}_%s.json"%d) if l.startswith("{")][0]
        print(d, "%.4e"%j["value"], "e2e %.4e"%j["e2e"]["value"], "ms/step %.2f"%j["ms_per_step"], j.get("clocks"))
    except Exception as e:
        print(d, "failed", e)
PY
tail -4 $O/dropin_${TAG}_csp.txt
