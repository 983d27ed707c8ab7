#!/usr/bin/env python3
"""Generates tests/golden/full_decks.json from the UNMODIFIED reference omp3 kernel set: the
BASELINE.json decks at FULL size (4000 x 4000 mesh, 1e6 / 1e7 particles).

Run in the build container, where /root/reference exists (oracle/Makefile compiles
omp3/neutral.c in place into oracle/_ref/libneutral_omp3.so):

    python tests/golden/make_golden_full.py [deck[@nparticles] ...]   # ~10 CPU-minutes on 8 cores

``deck@nparticles`` records the same deck with another particle count under that key: the
weak-scaled banks bench.py transports on N GPUs (csp@2000000, csp@4000000, csp@8000000, the
same for split) and the scaled-up split deck of BASELINE.json (split@100000000, ~40 minutes
and 24 GB of host memory on 8 cores).

Per deck it records, straight from the reference library:
  * per-timestep (facets, collisions)                                   [exact]
  * sha256 of each of the 11 particle fields of the final bank          [bit-exact]
  * sha256 of the injected bank's fields                                [bit-exact]
  * the sum of the final tally and a 64 x 64 block-sum image of it      [1e-9 relative]
The atomic summation order of the tally differs between runs of the reference itself (OpenMP
atomics), which is why the tally is pinned through sums and not through a hash.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from neutral_b200.bank import ALL_FIELDS, HostBank  # noqa: E402
from neutral_b200.decks import build_problem  # noqa: E402
from oracle.oracle import ReferenceOmp3  # noqa: E402

DECKS = ["split", "csp", "stream", "scatter"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_decks.json")


def field_hashes(bank: HostBank):
    return {k: hashlib.sha256(np.ascontiguousarray(bank.arrays[k]).tobytes()).hexdigest()
            for k in ALL_FIELDS}


def block_sums(tally, nx, ny, blocks=64):
    t = tally.reshape(ny, nx)
    by, bx = ny // blocks, nx // blocks
    return t[:by * blocks, :bx * blocks].reshape(blocks, by, blocks, bx).sum(axis=(1, 3))


def main():
    decks = sys.argv[1:] or DECKS
    ref = ReferenceOmp3()
    result = json.load(open(OUT)) if os.path.exists(OUT) else {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    for name in decks:
        base, _, count = name.partition("@")
        prob = build_problem(base, nparticles=int(count)) if count else build_problem(base)
        d = prob.deck
        t0 = time.time()
        aos = ref.inject(prob)
        inject = field_hashes(HostBank.from_aos(aos))
        tally = np.zeros(d.nx * d.ny)
        counts = []
        saved = os.dup(1)
        os.dup2(devnull, 1)  # the reference prints "Particles N" every timestep
        try:
            for tt in range(1, d.iterations + 1):
                counts.append(list(ref.step(prob, aos, tt, tally)))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        bank = HostBank.from_aos(aos)
        result[name] = {
            "nparticles": d.nparticles, "mesh": [d.nx, d.ny], "iterations": d.iterations,
            "counts": counts, "inject_hashes": inject, "final_hashes": field_hashes(bank),
            "live": int(np.count_nonzero(bank.dead == 0)),
            "tally_sum": float(np.sum(tally)),
            "tally_block_sums": block_sums(tally, d.nx, d.ny).ravel().tolist(),
        }
        print(f"{name}: {counts[-1]} live={result[name]['live']} "
              f"tally_sum={result[name]['tally_sum']:.15e}  ({time.time() - t0:.0f} s)",
              flush=True)
        with open(OUT, "w") as f:
            json.dump(result, f)


if __name__ == "__main__":
    main()
