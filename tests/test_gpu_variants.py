"""CUDA path vs the oracle port on the problem variants of tests/variants.py (branches the
reference decks never take): every particle field bit-exact after every timestep, exact
per-particle and aggregate event counts, tally within 1e-10 relative per cell."""
import numpy as np
import pytest

from neutral_b200.host import NB200_BAD_OPTION, Simulation
from test_gpu_parity import DEFAULTS, MODES, tally_close
from variants import VARIANTS

pytestmark = pytest.mark.gpu

CONFIGS = ["pipeline", "pipeline-ieee-div", "direct"]


@pytest.mark.parametrize("config", CONFIGS)
@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_matches_oracle_every_step(gpu_lib, port, name, config):
    for k, v in MODES[config].items():
        assert gpu_lib.nb200_set_option(k.encode(), v) != NB200_BAD_OPTION
    try:
        prob = VARIANTS[name]()
        d = prob.deck
        sim = Simulation(prob)
        sim.inject()
        bank = port.inject(prob)
        assert sum(sim.bank_to_host().bit_equal(bank).values()) == 0, "inject_particles differs"
        tally = np.zeros(d.nx * d.ny)
        ctr = np.zeros((3, len(bank)), dtype=np.uint64)
        for tt in range(1, d.iterations + 1):
            pf, pc, pp = port.step(prob, bank, tt, tally, counters=ctr)
            r = sim.step(tt)
            assert (r.facets, r.collisions, r.processed) == (pf, pc, pp), f"step {tt}"
            diff = sim.bank_to_host().bit_equal(bank)
            assert sum(diff.values()) == 0, f"step {tt}: {diff}"
            assert np.array_equal(sim.counters_to_host(), ctr), f"step {tt}: per-particle counts"
            assert tally_close(sim.tally_to_host(), tally), f"step {tt}: tally"
        sim.free()
    finally:
        for k, v in DEFAULTS.items():
            gpu_lib.nb200_set_option(k.encode(), v)
