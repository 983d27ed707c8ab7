#!/usr/bin/env python3
"""Per-timestep breakdown of a deck on one GPU: events by kind, history-kernel and sort-phase
milliseconds, and the L2-atomic ceiling of the step (tools/microbench/red_rate.cu) next to it.

    python tools/step_breakdown.py [deck] [--opts k=v,...] [--repeat R]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neutral_b200.decks import build_problem, load_deck  # noqa: E402
from neutral_b200.host import Simulation, load_library  # noqa: E402

RED_PEAK = 1.97e11

ap = argparse.ArgumentParser()
ap.add_argument("deck", nargs="?", default="csp")
ap.add_argument("--opts", default="")
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--particles", type=int, default=0)
args = ap.parse_args()

lib = load_library(build=False)
lib.nb200_set_option(b"print", 0)
for kv in args.opts.split(","):
    if kv:
        k, v = kv.split("=")
        lib.nb200_set_option(k.encode(), int(v))
deck = load_deck(args.deck)
prob = build_problem(deck, nparticles=args.particles or deck.nparticles)
sim = Simulation(prob, per_particle_counters=False)
for rep in range(args.repeat):
    sim.inject()
    sim.tally.zero()
    rows = sim.run()
    if rep < args.repeat - 1:
        continue
    print(f"{args.deck} opts={args.opts or 'default'}")
    print(" tt   processed      facets  collisions   census  deaths  hist_ms  sort_ms  "
          "atomic_floor_ms  Mev/s")
    tot_ms = 0.0
    tot_ev = 0
    for tt, r in enumerate(rows, 1):
        floor_ms = (r.facets + r.census + r.deaths) / RED_PEAK * 1e3
        ms = (r.kernel_ns + r.sort_ns) / 1e6
        tot_ms += ms
        tot_ev += r.events
        print(f"{tt:3d} {r.processed:11d} {r.facets:11d} {r.collisions:11d} {r.census:8d} "
              f"{r.deaths:7d} {r.kernel_ns / 1e6:8.3f} {r.sort_ns / 1e6:8.3f} {floor_ms:12.3f} "
              f"{r.events / max(ms, 1e-9) / 1e3:10.0f}")
    print(f"total {tot_ms:.3f} ms  {tot_ev / tot_ms / 1e3:.0f} Mev/s")
