"""Python handles on the two CPU checkers (TEST INFRASTRUCTURE -- never imported by
``neutral_b200``):

* :class:`OraclePort` -- ``oracle/neutral_oracle.c``, our plain-C restatement of the omp3 hot
  path with per-particle event counters and a global-pid offset;
* :class:`ReferenceOmp3` -- the UNMODIFIED reference ``omp3/neutral.c`` compiled by
  ``oracle/Makefile`` into ``oracle/_ref/libneutral_omp3.so`` (built where
  ``/root/reference`` exists; travels to the GPU box as a binary).

Both take the host arrays of :class:`neutral_b200.decks.Problem` and
:class:`neutral_b200.bank.HostBank`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

from neutral_b200.bank import PARTICLE_AOS, HostBank, ParticleSoA
from neutral_b200.decks import Problem

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libneutral_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libneutral_omp3.so")
REF_RUN_DIR = os.path.join(HERE, "_ref", "run", "neutral")
REF_EXE = os.path.join(REF_RUN_DIR, "neutral.omp3")

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


def build(ref: bool = True) -> None:
    """Compiles the port (needs gcc only) and, when the reference tree is present, the
    reference omp3 library and driver."""
    subprocess.run(["make", "-s", "-C", HERE, "port"] + (["ref"] if ref else []), check=True)


def _ptr(a: np.ndarray, ty=_dp):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ty)


class OraclePort:
    def __init__(self):
        if not os.path.exists(PORT_SO) or \
                os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "neutral_oracle.c")):
            build(ref=False)
        self.lib = C.CDLL(PORT_SO)
        L = self.lib
        L.oracle_threefry2x64_20.argtypes = [C.c_uint64] * 4 + [_u64p]
        L.oracle_random_pair.argtypes = [C.c_uint64] * 3 + [_dp, _dp]
        L.oracle_cs_index.argtypes = [_dp, C.c_int, C.c_double]
        L.oracle_cs_index.restype = C.c_int
        L.oracle_cs_lookup.argtypes = [_dp, _dp, C.c_int, C.c_double]
        L.oracle_cs_lookup.restype = C.c_double
        L.oracle_inject.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_double] * 5 + \
            [_dp, _dp, C.c_double, C.POINTER(ParticleSoA)]
        L.oracle_transport_step.argtypes = [
            C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int,
            C.POINTER(ParticleSoA), _dp, _dp, _dp, _dp, _dp, C.c_int, _dp, _dp, C.c_int,
            _dp, _u64p, _u64p, _u64p, _u64p]
        L.oracle_tally_sum.argtypes = [_dp, C.c_size_t]
        L.oracle_tally_sum.restype = C.c_double

    # -- known-answer level ---------------------------------------------------------
    def threefry(self, c0, c1, k0, k1) -> Tuple[int, int]:
        out = (C.c_uint64 * 2)()
        self.lib.oracle_threefry2x64_20(c0, c1, k0, k1, out)
        return int(out[0]), int(out[1])

    def random_pair(self, pkey, master_key, counter) -> Tuple[float, float]:
        a, b = C.c_double(), C.c_double()
        self.lib.oracle_random_pair(pkey, master_key, counter, C.byref(a), C.byref(b))
        return a.value, b.value

    def cs_index(self, keys: np.ndarray, energy: float) -> int:
        return self.lib.oracle_cs_index(_ptr(keys), len(keys), energy)

    def cs_lookup(self, keys, values, energy: float) -> float:
        return self.lib.oracle_cs_lookup(_ptr(keys), _ptr(values), len(keys), energy)

    # -- deck level -------------------------------------------------------------------
    def inject(self, prob: Problem, pid0: int = 0, count: Optional[int] = None) -> HostBank:
        d, s = prob.deck, prob.source
        count = d.nparticles - pid0 if count is None else count
        bank = HostBank.empty(count)
        st = bank.as_struct()
        self.lib.oracle_inject(pid0, count, d.nx, d.ny, s.left, s.bottom, s.width, s.height,
                               d.dt, _ptr(prob.edgex), _ptr(prob.edgey), d.initial_energy,
                               C.byref(st))
        return bank

    def step(self, prob: Problem, bank: HostBank, master_key: int, tally: np.ndarray,
             pid0: int = 0, counters=None, ntotal: Optional[int] = None):
        """One timestep in place; returns (facets, collisions, processed). ``counters`` is
        an optional (3, n) uint64 array accumulating per-particle facets/collisions/census."""
        d = prob.deck
        st = bank.as_struct()
        totals = (C.c_uint64 * 3)()
        if counters is not None:
            assert counters.shape == (3, len(bank)) and counters.dtype == np.uint64
            cf, cc, cz = (_ptr(counters[i], _u64p) for i in range(3))
        else:
            cf = cc = cz = None
        (sk, sv), (ak, av) = prob.cs_scatter, prob.cs_absorb
        self.lib.oracle_transport_step(
            d.nx, d.ny, master_key, d.dt, ntotal or d.nparticles, pid0, len(bank),
            C.byref(st), _ptr(prob.density), _ptr(prob.edgex), _ptr(prob.edgey),
            _ptr(sk), _ptr(sv), len(sk), _ptr(ak), _ptr(av), len(ak),
            _ptr(tally), cf, cc, cz, totals)
        return int(totals[0]), int(totals[1]), int(totals[2])


class _CrossSection(C.Structure):  # reference neutral_data.h:38-43
    _fields_ = [("keys", _dp), ("values", _dp), ("nentries", C.c_int)]


class ReferenceOmp3:
    """The reference's own three boundary functions (neutral_interface.h:11-36), AoS."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference"):
                build(ref=True)
            else:
                raise FileNotFoundError(f"{REF_SO} is missing and /root/reference is absent")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.inject_particles.restype = C.c_size_t
        L.inject_particles.argtypes = [C.c_int] * 5 + [C.c_double] * 4 + [C.c_int, C.c_int,
                                      C.c_double, _dp, _dp, C.c_double,
                                      C.POINTER(C.c_void_p)]
        L.solve_transport_2d.restype = None
        L.solve_transport_2d.argtypes = [
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
            C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
            _dp, _dp, _dp, _dp, _dp, C.POINTER(_CrossSection), C.POINTER(_CrossSection),
            _dp, _u64p, _u64p, _u64p, _u64p, _u64p]
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO) or os.path.isdir("/root/reference")

    def inject(self, prob: Problem) -> np.ndarray:
        """AoS bank (numpy structured array) as the reference initialises it."""
        d, s = prob.deck, prob.source
        out = C.c_void_p()
        self.lib.inject_particles(d.nparticles, d.nx, d.nx, d.ny, 0, s.left, s.bottom,
                                  s.width, s.height, 0, 0, d.dt, _ptr(prob.edgex),
                                  _ptr(prob.edgey), d.initial_energy, C.byref(out))
        buf = (C.c_char * (PARTICLE_AOS.itemsize * d.nparticles)).from_address(out.value)
        aos = np.frombuffer(buf, dtype=PARTICLE_AOS).copy()
        self._libc.free(out)
        return aos

    def step(self, prob: Problem, aos: np.ndarray, master_key: int, tally: np.ndarray):
        """solve_transport_2d on a host AoS bank, in place; returns (facets, collisions)."""
        d = prob.deck
        assert aos.dtype == PARTICLE_AOS and aos.flags["C_CONTIGUOUS"]
        (sk, sv), (ak, av) = prob.cs_scatter, prob.cs_absorb
        cs_s = _CrossSection(_ptr(sk), _ptr(sv), len(sk))
        cs_a = _CrossSection(_ptr(ak), _ptr(av), len(ak))
        nlocal = C.c_int(len(aos))
        neigh = (C.c_int * 6)(*([-1] * 6))
        facets, colls = C.c_uint64(0), C.c_uint64(0)
        dx = np.full(d.nx + 1, d.width / d.nx)
        dy = np.full(d.ny + 1, d.height / d.ny)
        self.lib.solve_transport_2d(
            d.nx, d.ny, d.nx, d.ny, master_key, 0, 0, 0, d.dt, d.nparticles,
            C.byref(nlocal), neigh, aos.ctypes.data_as(C.c_void_p), _ptr(prob.density),
            _ptr(prob.edgex), _ptr(prob.edgey), _ptr(dx), _ptr(dy), C.byref(cs_s),
            C.byref(cs_a), _ptr(tally), None, None, None, C.byref(facets), C.byref(colls))
        return int(facets.value), int(colls.value)
