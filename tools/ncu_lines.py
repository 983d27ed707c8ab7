#!/usr/bin/env python3
"""Per-source-line instruction counts of one kernel from an ncu report.

    tools/ncu_lines.py rep.ncu-rep <kernel-mangled-substring> [lib.so] [kernel-id] [topN]

`ncu --page source --csv` lists SASS only; this joins its rows (in instruction order) with
the line table of the same function (`nvdisasm -g` on the cubin extracted from the library)
and prints, per source line, the warp instructions executed, the share of the kernel's
total, and the average active threads - the figures behind DESIGN.md's cost breakdown of an
event. Lines are reported twice: innermost location (the helper a cost sits in) and outermost
location (the statement of the kernel that pulled the helper in).
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep = sys.argv[1]
kname = sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "neutral_b200/libneutral_b200.so"
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 45

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout
blocks, cur = [], []
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    else:
        cur.append(line)
if cur:
    blocks.append(cur)
rows = list(csv.reader(io.StringIO("\n".join(blocks[which][1:]))))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
sass = rows[1:]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
locs = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True,
                         text=True).stdout
    if kname not in txt:
        continue
    locs = []
    inside = False
    cur_loc = (("?", 0), ("?", 0))
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            inside = kname in ln
            continue
        if not inside:
            continue
        m = re.findall(r'File "([^"]+)", line (\d+)', ln)
        if ln.lstrip().startswith("//## File") and m:
            inner = (os.path.basename(m[0][0]), int(m[0][1]))
            outer = (os.path.basename(m[-1][0]), int(m[-1][1]))
            cur_loc = (inner, outer)
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
            locs.append(cur_loc)
    if locs:
        break
if not locs:
    sys.exit(f"kernel {kname} not found in {lib}")
if len(locs) != len(sass):
    print(f"warning: {len(locs)} instructions in the cubin vs {len(sass)} in the report",
          file=sys.stderr)

tot_w = sum(int(r[ix["Instructions Executed"]]) for r in sass)
tot_t = sum(int(r[ix["Thread Instructions Executed"]]) for r in sass)
print(f"kernel {kname}: {tot_w} warp instructions, {tot_t} thread instructions, "
      f"efficiency {tot_t / 32 / max(tot_w, 1):.3f}")
for label, pick in (("innermost", 0), ("outermost", 1)):
    agg = defaultdict(lambda: [0, 0, 0])
    for loc, r in zip(locs, sass):
        a = agg[loc[pick]]
        a[0] += int(r[ix["Instructions Executed"]])
        a[1] += int(r[ix["Thread Instructions Executed"]])
        a[2] += int(r[ix["# Samples"]])
    print(f"-- by {label} source line: warp-instr share, thread efficiency, stall-sample share")
    tot_s = sum(v[2] for v in agg.values())
    for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {loc[0]:>16s}:{loc[1]:<4d} {100 * v[0] / tot_w:6.2f}%  eff {v[1] / 32 / max(v[0], 1):.2f}"
              f"  samples {100 * v[2] / max(tot_s, 1):5.1f}%")
