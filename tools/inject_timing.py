#!/usr/bin/env python3
"""inject_particles: device kernel vs host loop + upload, wall-clock per call.
    python tools/inject_timing.py [nparticles]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neutral_b200.decks import build_problem, load_deck  # noqa: E402
from neutral_b200.host import Simulation, load_library  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
lib = load_library(build=False)
prob = build_problem(load_deck("split"), nparticles=n)
sim = Simulation(prob, per_particle_counters=False)
for mode, name in ((1, "device"), (0, "host+upload")):
    lib.nb200_set_option(b"device_inject", mode)
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        sim.inject()
        best = min(best, time.perf_counter() - t0)
    print(f"inject_particles({n}) {name:12s} {best * 1e3:9.2f} ms  "
          f"(OMP threads: {os.environ.get('OMP_NUM_THREADS', 'all')})")
lib.nb200_set_option(b"device_inject", 1)
