#!/usr/bin/env python3
"""Timeline of one history-kernel launch, warp by warp (measurement build only):

    NB200_DEFINES=-DNB_TRACE_WARPS NB200_LIB=libneutral_b200.trace.so python -m neutral_b200.build
    NB200_LIB=libneutral_b200.trace.so python tools/warp_trace.py csp --step 5

Every warp of k_history records start, end, SM and its event counts; this prints, for the
chosen timestep, when the collider warps (any collision) and the streamer warps start and end,
how many of each are resident over time, and what the tail of the launch consists of.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neutral_b200.decks import build_problem, load_deck  # noqa: E402
from neutral_b200.host import DeviceArray, Simulation, load_library  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("deck", nargs="?", default="csp")
ap.add_argument("--step", type=int, default=5)
ap.add_argument("--opts", default="")
ap.add_argument("--bins", type=int, default=24)
args = ap.parse_args()

lib = load_library(build=False)
lib.nb200_set_option(b"print", 0)
for kv in args.opts.split(","):
    if kv:
        k, v = kv.split("=")
        lib.nb200_set_option(k.encode(), int(v))
deck = load_deck(args.deck)
prob = build_problem(deck)
sim = Simulation(prob, per_particle_counters=False)
sim.inject()
nwarps = (prob.deck.nparticles + 127) // 128 * 4
trace = DeviceArray(6 * nwarps, np.uint64)
for tt in range(1, args.step):
    sim.step(tt)
lib.nb200_debug_set_warp_trace.argtypes = [C.c_void_p]
assert lib.nb200_debug_set_warp_trace(trace.ptr) == 0
r = sim.step(args.step)
lib.nb200_debug_set_warp_trace(None)
t = trace.download().reshape(-1, 6)
t = t[t[:, 1] > 0]
t0 = t[:, 0].min()
start, end = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6  # ms
coll = t[:, 4] > 0
live = t[:, 5] > 0
print(f"{args.deck} step {args.step}: kernel {r.kernel_ns / 1e6:.3f} ms, {len(t)} warps recorded, "
      f"{int(coll.sum())} collider warps, {int((~coll & live).sum())} streamer warps")
late = coll & (start > 0.05)  # colliders dispatched behind streamers (option stagger_at)
groups = [("collider", coll), ("streamer", ~coll & live)]
if late.any():
    groups = [("collider@0", coll & ~late), ("collider late", late), ("streamer", ~coll & live)]
    sm_late = np.bincount(t[late, 2].astype(int), minlength=148)
    print(f"  late collider warps per SM: min {sm_late.min()} median {int(np.median(sm_late))} "
          f"max {sm_late.max()} (SMs without any: {int((sm_late == 0).sum())})")
for name, m in groups:
    if m.any():
        d = end[m] - start[m]
        print(f"  {name:9s} start {start[m].min():.3f}..{start[m].max():.3f} ms  end "
              f"{end[m].min():.3f}..{end[m].max():.3f} ms  lifetime mean {d.mean():.3f} "
              f"p50 {np.median(d):.3f} p99 {np.percentile(d, 99):.3f} max {d.max():.3f} ms  "
              f"events/warp {int((t[m, 3] + t[m, 4]).mean())}")
span = end.max()
print(f"  launch span {span:.3f} ms; resident warps per SM over time (collider / streamer), "
      f"{args.bins} bins:")
edges = np.linspace(0.0, span, args.bins + 1)
for i in range(args.bins):
    mid = 0.5 * (edges[i] + edges[i + 1])
    rc = int(((start <= mid) & (end > mid) & coll).sum())
    rs = int(((start <= mid) & (end > mid) & ~coll).sum())
    print(f"    {mid:6.3f} ms  {rc / 148:5.1f} / {rs / 148:5.1f}")
