"""Problem variants the reference's own decks never exercise (test infrastructure).

The four reference decks share three properties that hide whole branches of the hot path:
both cross-section tables are the same file (so ``p_absorb == 0.5`` exactly and the absorb
lookup can reuse the scatter lookup), the mesh is square and uniform, and
every density box is aligned to a multiple of 16 cells. The variants below break each of
them; both CPU checkers (the oracle port and the unmodified reference library) and the CUDA
path must still agree bit for bit on them.
"""
from __future__ import annotations

import numpy as np

from neutral_b200.decks import (Deck, Problem, build_problem, cross_section_table,
                                density_field, source_box)


def absorb_scaled(base: str = "mixed_small", factor: float = 0.25) -> Problem:
    """Same energy grid, capture = factor x elastic: p_absorb = factor / (1 + factor)."""
    prob = build_problem(base)
    keys, vals = prob.cs_scatter
    prob.cs_absorb = (keys.copy(), np.ascontiguousarray(vals * factor))
    return prob


def absorb_own_grid(base: str = "mixed_small", nentries: int = 7001) -> Problem:
    """Capture table on its own (geometric) energy grid with its own shape: the two lookups
    bracket different intervals and p_absorb varies with the energy."""
    prob = build_problem(base)
    keys = np.geomspace(5.0e-3, 3.0e8, nentries)
    vals = 40.0 + 900.0 / (1.0 + np.log10(keys / 5.0e-3))
    prob.cs_absorb = (np.ascontiguousarray(keys), np.ascontiguousarray(vals))
    return prob


def scatter_short_table(base: str = "csp_small", nentries: int = 257) -> Problem:
    """A coarse elastic table (every ~117th point of the reference's): few grid points per
    bucket of the staged index, long interpolation intervals."""
    prob = build_problem(base)
    keys, vals = cross_section_table()
    pick = np.unique(np.linspace(0, len(keys) - 1, nentries).astype(int))
    prob.cs_scatter = (np.ascontiguousarray(keys[pick]), np.ascontiguousarray(vals[pick]))
    return prob


def _custom(deck: Deck, edgex: np.ndarray, edgey: np.ndarray) -> Problem:
    return Problem(deck=deck, edgex=edgex, edgey=edgey,
                   density=density_field(deck, edgex, edgey),
                   source=source_box(deck, edgex, edgey),
                   cs_scatter=cross_section_table(), cs_absorb=cross_section_table())


def rect_stretched() -> Problem:
    """150 x 70 cells on a 2.0 x 0.7 box with smoothly stretched (non-uniform) edges, a thin
    medium with two denser inclusions whose borders fall inside 16-cell tiles."""
    deck = Deck(path="rect_stretched.params", nx=150, ny=70, dt=1.0e-7, iterations=3,
                nparticles=3000, initial_energy=2.0e5,
                source=(0.30, 0.20, 0.35, 0.55),
                problems=[(1.0e-30, 0.0, 0.0, 1.0, 1.0),
                          (2.0e0, 0.13, 0.11, 0.61, 0.73),
                          (9.0e0, 0.41, 0.37, 0.17, 0.29)],
                width=2.0, height=0.7)
    tx = np.arange(deck.nx + 1, dtype=np.float64) / deck.nx
    ty = np.arange(deck.ny + 1, dtype=np.float64) / deck.ny
    edgex = deck.width * (tx + 0.35 * tx * tx) / 1.35
    edgey = deck.height * (ty + 0.5 * ty * ty * ty) / 1.5
    edgex[-1], edgey[-1] = deck.width, deck.height
    return _custom(deck, np.ascontiguousarray(edgex), np.ascontiguousarray(edgey))


def multi_tile() -> Problem:
    """400 x 300 uniform cells: several coarse (256-cell) and fine (16-cell) tiles of the
    density maps, with boxes that straddle both kinds of tile border."""
    deck = Deck(path="multi_tile.params", nx=400, ny=300, dt=1.0e-7, iterations=3,
                nparticles=5000, initial_energy=1.0e5,
                source=(0.05, 0.05, 0.9, 0.9),
                problems=[(1.0e-30, 0.0, 0.0, 1.0, 1.0),
                          (1.5e0, 0.07, 0.09, 0.33, 0.41),
                          (4.0e0, 0.31, 0.28, 0.45, 0.37),
                          (1.0e-30, 0.47, 0.44, 0.09, 0.07),
                          (2.5e1, 0.655, 0.605, 0.02, 0.0234)])
    dx, dy = deck.width / deck.nx, deck.height / deck.ny
    return _custom(deck, dx * np.arange(deck.nx + 1, dtype=np.float64),
                   dy * np.arange(deck.ny + 1, dtype=np.float64))


VARIANTS = {
    "absorb_scaled": absorb_scaled,
    "absorb_scaled_csp": lambda: absorb_scaled("csp_small", 3.0),
    "absorb_own_grid": absorb_own_grid,
    "scatter_short_table": scatter_short_table,
    "rect_stretched": rect_stretched,
    "multi_tile": multi_tile,
}
