"""Problem decks: the host-side set-up that surrounds neutral's hot path.

This mirrors, for Python callers (tests, bench.py), what the reference's C driver does
before the first ``solve_transport_2d`` call:

* deck grammar -- ``name value`` and ``name k=v k=v ...`` lines with ``#`` comments
  (reference ``problems/csp.params:1-10``; reader = arch ``get_*_parameter``, restated in
  ``archlite/archlite.c``);
* the uniform mesh of ``initialise_mesh_2d`` (``main.c:65``; ``archlite/archlite.c``);
* the density field painted from the ``problem_<n>`` boxes (``main.c:66-68``);
* the source rectangle and local particle count (``neutral_data.c:39-95``);
* the cross-section tables (``neutral_data.c:123-170``), which are regenerated here from
  the formula of the reference's ``resonance.py:25-29,41-43`` -- the text they print is
  byte-identical to ``elastic_scatter.cs`` / ``capture.cs`` (md5 pinned below).

Everything is IEEE binary64 with the same operation order as the C code so that Python and
the C driver produce bit-identical inputs (checked in ``tests/test_decks.py``).
"""
from __future__ import annotations

import hashlib
import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBLEMS_DIR = os.path.join(REPO_ROOT, "problems")
ARCH_PARAMS = os.path.join(REPO_ROOT, "archlite", "arch.params")

#: md5 of the reference's elastic_scatter.cs and capture.cs (they are the same file).
CS_TABLE_MD5 = "6deb6261687eb2c528e9f4f1bff12793"
CS_TABLE_ENTRIES = 29999


# --------------------------------------------------------------------------- parsing --


def _lines(path: str):
    with open(path, "r") as f:
        for raw in f:
            line = raw.split("#", 1)[0].strip()
            if line:
                yield line


def _find(path: str, name: str) -> Optional[str]:
    """Text after the first line whose first token is ``name`` (None if absent)."""
    for line in _lines(path):
        parts = line.split(None, 1)
        if parts[0] == name:
            return parts[1] if len(parts) > 1 else ""
    return None


def get_double_parameter(name: str, path: str) -> float:
    rest = _find(path, name)
    if rest is None or not rest.split():
        raise KeyError(f"Parameter {name} was not found in {path}")
    return float(rest.split()[0])


def get_int_parameter(name: str, path: str) -> int:
    """Like the C reader's strtol: the leading integer of the value text."""
    rest = _find(path, name)
    m = re.match(r"\s*[+-]?\d+", rest or "")
    if m is None:
        raise KeyError(f"Parameter {name} was not found in {path}")
    return int(m.group(0))


def get_key_value_parameter(name: str, path: str) -> Optional[List[Tuple[str, float]]]:
    rest = _find(path, name)
    if rest is None:
        return None
    out = []
    for tok in rest.split():
        if "=" in tok:
            k, v = tok.split("=", 1)
            out.append((k, float(v)))
    return out


# ------------------------------------------------------------------------------ deck --


@dataclass
class Deck:
    """One problem deck plus the mesh extents that normally come from ``../arch.params``."""

    path: str
    nx: int
    ny: int
    dt: float
    iterations: int
    nparticles: int
    initial_energy: float
    source: Tuple[float, float, float, float]  # xpos ypos width height (fractions)
    problems: List[Tuple[float, float, float, float, float]] = field(default_factory=list)
    width: float = 1.0
    height: float = 1.0
    sim_end: float = 10.0

    @property
    def name(self) -> str:
        return os.path.splitext(os.path.basename(self.path))[0]

    def scaled(self, **changes) -> "Deck":
        """Copy with some parameters replaced (e.g. ``nparticles`` for a bounded sample)."""
        d = Deck(**{**self.__dict__})
        for k, v in changes.items():
            if not hasattr(d, k):
                raise AttributeError(k)
            setattr(d, k, v)
        return d


def find_deck(name_or_path: str) -> str:
    """Resolves 'csp', 'small/csp_small', 'problems/csp.params' or a real path."""
    cands = [name_or_path,
             os.path.join(REPO_ROOT, name_or_path),
             os.path.join(PROBLEMS_DIR, name_or_path),
             os.path.join(PROBLEMS_DIR, name_or_path + ".params"),
             os.path.join(PROBLEMS_DIR, "small", name_or_path + ".params")]
    for c in cands:
        if os.path.isfile(c):
            return c
    raise FileNotFoundError(f"no deck named {name_or_path!r}")


def load_deck(name_or_path: str, arch_params: str = ARCH_PARAMS) -> Deck:
    path = find_deck(name_or_path)
    src = get_key_value_parameter("source", path)
    if src is None or len(src) < 4:
        raise ValueError(f"Parameter file {path} did not contain a source entry.")
    src_vals = [v for _, v in src]
    problems = []
    pp = 0
    while True:
        kv = get_key_value_parameter(f"problem_{pp}", path)
        if kv is None:
            break
        if len(kv) < 5:
            raise ValueError(f"problem_{pp} of {path} needs density and a box")
        vals = [v for _, v in kv]
        rho = dict(kv).get("density", vals[0])
        problems.append((rho, vals[-4], vals[-3], vals[-2], vals[-1]))
        pp += 1
    return Deck(
        path=path,
        nx=get_int_parameter("nx", path),
        ny=get_int_parameter("ny", path),
        dt=get_double_parameter("dt", path),
        iterations=get_int_parameter("iterations", path),
        nparticles=get_int_parameter("nparticles", path),
        initial_energy=get_double_parameter("initial_energy", path),
        source=tuple(src_vals[-4:]),
        problems=problems,
        width=get_double_parameter("width", arch_params),
        height=get_double_parameter("height", arch_params),
        sim_end=get_double_parameter("sim_end", arch_params),
    )


# ------------------------------------------------------------------------------ mesh --


def mesh_edges(deck: Deck) -> Tuple[np.ndarray, np.ndarray]:
    """nx+1 / ny+1 edge coordinates: edge i = (width / nx) * i  (pad = 0, offsets = 0)."""
    dx = deck.width / float(deck.nx)
    dy = deck.height / float(deck.ny)
    edgex = dx * np.arange(deck.nx + 1, dtype=np.float64)
    edgey = dy * np.arange(deck.ny + 1, dtype=np.float64)
    return edgex, edgey


def density_field(deck: Deck, edgex: np.ndarray, edgey: np.ndarray) -> np.ndarray:
    """(ny, nx) density; a cell belongs to a box when its lower-left edge is in [pos, pos+size)."""
    rho = np.zeros((deck.ny, deck.nx), dtype=np.float64)
    ex = edgex[: deck.nx]
    ey = edgey[: deck.ny]
    for value, fx, fy, fw, fh in deck.problems:
        xpos = fx * deck.width
        ypos = fy * deck.height
        xend = xpos + fw * deck.width
        yend = ypos + fh * deck.height
        mx = (ex >= xpos) & (ex < xend)
        my = (ey >= ypos) & (ey < yend)
        rho[np.ix_(my, mx)] = value
    return rho


@dataclass
class SourceBox:
    left: float
    bottom: float
    width: float
    height: float
    nlocal_particles: int


def source_box(deck: Deck, edgex: np.ndarray, edgey: np.ndarray) -> SourceBox:
    """The arguments ``initialise_neutral_data`` hands to ``inject_particles``
    (reference ``neutral_data.c:39-95,109-114``), for the single-rank mesh."""
    sx = deck.source[0] * deck.width
    sy = deck.source[1] * deck.height
    sw = deck.source[2] * deck.width
    sh = deck.source[3] * deck.height
    x0, x1 = float(edgex[0]), float(edgex[deck.nx])
    y0, y1 = float(edgey[0]), float(edgey[deck.ny])
    left = max(0.0, sx - x0)
    bottom = max(0.0, sy - y0)
    right = max(0.0, x1 - (sx + sw))
    top = max(0.0, y1 - (sy + sh))
    width = max(0.0, (x1 - x0) - (right + left))
    height = max(0.0, (y1 - y0) - (top + bottom))
    nlocal_real = deck.nparticles * (width * height) / (sw * sh)
    return SourceBox(left, bottom, width, height, int(nlocal_real + 0.5))


# -------------------------------------------------------------------- cross sections --


def cross_section_text() -> bytes:
    """The text of the reference's dummy resonance table (both .cs files are this text)."""
    t = np.linspace(0, 1, num=CS_TABLE_ENTRIES + 1)
    energy = 10.0e7 * np.power(t, 4) + 10.0e-3
    cs = 1.0e3 * t + 1.0
    n = len(energy)
    txt = "".join("%.12e %.12e\n" % (energy[rr], cs[n - rr]) for rr in range(1, n)).encode()
    return txt


_CS_CACHE: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}


def cross_section_table(path: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
    """(keys, values) as the C reader parses them (``fscanf("%lf")`` == Python float())."""
    key = path or "<generated>"
    if key not in _CS_CACHE:
        if path is None:
            txt = cross_section_text()
            if hashlib.md5(txt).hexdigest() != CS_TABLE_MD5:
                raise RuntimeError("regenerated cross-section table does not match the "
                                   "reference table's md5")
        else:
            with open(path, "rb") as f:
                txt = f.read()
        arr = np.array(txt.split(), dtype=np.float64).reshape(-1, 2)
        _CS_CACHE[key] = (np.ascontiguousarray(arr[:, 0]), np.ascontiguousarray(arr[:, 1]))
    k, v = _CS_CACHE[key]
    return k.copy(), v.copy()


def write_cross_section_files(directory: str) -> None:
    """Drops elastic_scatter.cs and capture.cs (what the C driver opens by relative path)."""
    txt = cross_section_text()
    for name in ("elastic_scatter.cs", "capture.cs"):
        with open(os.path.join(directory, name), "wb") as f:
            f.write(txt)


# -------------------------------------------------------------------------- problem ---


@dataclass
class Problem:
    """Everything ``solve_transport_2d`` reads, as host arrays."""

    deck: Deck
    edgex: np.ndarray
    edgey: np.ndarray
    density: np.ndarray
    source: SourceBox
    cs_scatter: Tuple[np.ndarray, np.ndarray]
    cs_absorb: Tuple[np.ndarray, np.ndarray]


def build_problem(deck_or_name, **scale) -> Problem:
    deck = deck_or_name if isinstance(deck_or_name, Deck) else load_deck(deck_or_name)
    if scale:
        deck = deck.scaled(**scale)
    edgex, edgey = mesh_edges(deck)
    return Problem(
        deck=deck,
        edgex=edgex,
        edgey=edgey,
        density=density_field(deck, edgex, edgey),
        source=source_box(deck, edgex, edgey),
        cs_scatter=cross_section_table(),
        cs_absorb=cross_section_table(),
    )


def shard_range(ntotal: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Contiguous particle range of one GPU: the same split the reference uses for its
    OpenMP threads (``omp3/neutral.c:64-74``). Returns (first global pid, count)."""
    per = ntotal // nranks
    rem = ntotal % nranks
    return rank * per + min(rank, rem), per + (1 if rank < rem else 0)
