"""The oracle port against the UNMODIFIED reference library on problems the reference decks do
not cover (tests/variants.py): distinct capture/elastic tables (values only, and grid +
values), a coarse table, a rectangular non-uniform mesh, density boxes off the tile grid."""
import numpy as np
import pytest

from neutral_b200.bank import HostBank
from variants import VARIANTS


@pytest.mark.parametrize("name", list(VARIANTS))
def test_port_matches_reference_build_on_variant(port, ref, name):
    prob = VARIANTS[name]()
    d = prob.deck
    aos = ref.inject(prob)
    bank = port.inject(prob)
    assert sum(HostBank.from_aos(aos).bit_equal(bank).values()) == 0, "inject"
    t_ref, t_port = np.zeros(d.nx * d.ny), np.zeros(d.nx * d.ny)
    events = 0
    for tt in range(1, d.iterations + 1):
        f, c = ref.step(prob, aos, tt, t_ref)
        pf, pc, _ = port.step(prob, bank, tt, t_port)
        assert (f, c) == (pf, pc), f"step {tt}"
        assert sum(HostBank.from_aos(aos).bit_equal(bank).values()) == 0, f"step {tt}"
        events += f + c
    assert np.all(np.abs(t_ref - t_port) <= 1e-12 * np.maximum(np.abs(t_ref), np.abs(t_port)))
    # the variant must actually exercise both event kinds
    assert events > 10 * d.nparticles


def test_variants_break_the_reference_decks_symmetries():
    p = VARIANTS["absorb_scaled"]()
    assert np.array_equal(p.cs_scatter[0], p.cs_absorb[0])
    assert not np.array_equal(p.cs_scatter[1], p.cs_absorb[1])
    p = VARIANTS["absorb_own_grid"]()
    assert len(p.cs_scatter[0]) != len(p.cs_absorb[0])
    assert np.all(np.diff(p.cs_absorb[0]) > 0)
    p = VARIANTS["rect_stretched"]()
    assert p.deck.nx != p.deck.ny
    assert np.ptp(np.diff(p.edgex)) > 1e-3 and np.all(np.diff(p.edgex) > 0)
    assert np.all(np.diff(p.edgey) > 0)
    for name in ("rect_stretched", "multi_tile"):
        rho = VARIANTS[name]().density
        # at least one 16x16 tile holds more than one density value
        ny, nx = rho.shape
        mixed = any(np.unique(rho[y:y + 16, x:x + 16]).size > 1
                    for y in range(0, ny, 16) for x in range(0, nx, 16))
        assert mixed, name
