#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference omp3 kernel set.

Run in the build container, where /root/reference exists (oracle/Makefile compiles
omp3/neutral.c in place into oracle/_ref/libneutral_omp3.so):

    OMP_NUM_THREADS=1 python tests/golden/make_golden.py

For every small parity deck it records, straight from the reference library:
  * per-timestep aggregate (facets, collisions)                        [exact]
  * sha256 of every particle field after step 1 and after the last step [bit-exact]
  * the first 256 particles of the final bank, all fields              [bit-exact]
  * the final energy-deposition tally                                   [1e-10 per cell]
and, from the oracle port (which has per-particle counters the reference lacks) after it
has been verified bit-equal to the reference on that deck:
  * cumulative per-particle facet / collision / census counts          [exact]
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from neutral_b200.bank import ALL_FIELDS, HostBank  # noqa: E402
from neutral_b200.decks import build_problem  # noqa: E402
from oracle.oracle import OraclePort, ReferenceOmp3  # noqa: E402

DECKS = ["csp_small", "split_small", "scatter_small", "stream_small", "mixed_small"]
SAMPLE = 256


def field_hashes(bank: HostBank):
    return np.array([hashlib.sha256(np.ascontiguousarray(bank.arrays[k]).tobytes()).hexdigest()
                     for k in ALL_FIELDS])


def main():
    ref, port = ReferenceOmp3(), OraclePort()
    here = os.path.dirname(os.path.abspath(__file__))
    for name in DECKS:
        prob = build_problem(name)
        d = prob.deck
        aos = ref.inject(prob)
        bank = port.inject(prob)
        assert sum(HostBank.from_aos(aos).bit_equal(bank).values()) == 0
        inject_hashes = field_hashes(bank)
        t_ref = np.zeros(d.nx * d.ny)
        t_port = np.zeros(d.nx * d.ny)
        counters = np.zeros((3, len(bank)), dtype=np.uint64)
        counts = []
        step1_hashes = None
        for tt in range(1, d.iterations + 1):
            f, c = ref.step(prob, aos, tt, t_ref)
            pf, pc, pp = port.step(prob, bank, tt, t_port, counters=counters)
            assert (f, c) == (pf, pc), (name, tt)
            assert sum(HostBank.from_aos(aos).bit_equal(bank).values()) == 0, (name, tt)
            counts.append((f, c, pp))
            if tt == 1:
                step1_hashes = field_hashes(HostBank.from_aos(aos))
        final = HostBank.from_aos(aos)
        out = os.path.join(here, f"{name}.npz")
        np.savez_compressed(
            out,
            fields=np.array(ALL_FIELDS),
            counts=np.array(counts, dtype=np.uint64),
            inject_hashes=inject_hashes,
            step1_hashes=step1_hashes,
            final_hashes=field_hashes(final),
            sample=final.to_aos()[:SAMPLE],
            tally=t_ref,
            tally_sum=np.float64(t_ref.sum()),
            particle_counters=counters,
            nparticles=np.int64(d.nparticles),
        )
        print(f"{name}: {d.iterations} steps, counts[-1]={counts[-1]}, "
              f"{os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
