#!/usr/bin/env python3
"""Top stall lines of an ncu source page: tools/ncu_hot.py rep.ncu-rep [topN] [kernel-id]
Prints the SASS instructions with the most warp-stall samples, with their dominant stall
reason, plus the executed-instruction totals of the hot loop."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] , capture_output=True, text=True).stdout
# several kernels may follow one another; split on "Kernel Name" rows
blocks, cur = [], []
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [line]
    else:
        cur.append(line)
if cur: blocks.append(cur)
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
b = blocks[which]
print(b[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[1:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("total samples", tot, "warp-instructions", tot_inst)
agg = {}
for r in data:
    for h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
print("stall mix:", ", ".join(f"{k[6:]}={v*100//max(tot,1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]])
    reasons = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {s*100/tot:5.1f}% exec={int(r[ix['Instructions Executed']]):>12d} thr={r[ix['Avg. Threads Executed']]:>5s} {reasons[0][1]}:{reasons[0][0]} {reasons[1][1]}:{reasons[1][0]}  | {r[ix['Source']].strip()}")
