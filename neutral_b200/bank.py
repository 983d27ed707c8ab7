"""Host-side views of a particle bank.

Two layouts cross neutral's plugin boundary (reference ``neutral_data.h``):

* AoS, 80 bytes per particle (``neutral_data.h:66-79``) -- what the omp3 kernel set and the
  host-buffer flavour of the b200 C-ABI exchange (:data:`PARTICLE_AOS`);
* SoA, a struct of 11 array pointers (``neutral_data.h:48-61``, ``-DSoA``) -- what the GPU
  kernel sets and the device-resident flavour of the b200 C-ABI use
  (:class:`ParticleSoA`, ctypes).

:class:`HostBank` holds the 11 arrays as numpy arrays and converts between the two.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict

import numpy as np

F64_FIELDS = ("x", "y", "omega_x", "omega_y", "energy", "weight", "dt_to_census",
              "mfp_to_collision")
I32_FIELDS = ("cellx", "celly", "dead")
ALL_FIELDS = F64_FIELDS + I32_FIELDS

#: numpy dtype of the reference's AoS ``Particle`` (8 doubles, 3 ints, 4 bytes of padding).
PARTICLE_AOS = np.dtype(
    {"names": list(ALL_FIELDS),
     "formats": ["<f8"] * 8 + ["<i4"] * 3,
     "offsets": [0, 8, 16, 24, 32, 40, 48, 56, 64, 68, 72],
     "itemsize": 80})


class ParticleSoA(C.Structure):
    """ctypes mirror of the reference's ``-DSoA`` ``Particle`` (member order matters)."""

    _fields_ = [(n, C.POINTER(C.c_double)) for n in F64_FIELDS] + \
               [(n, C.POINTER(C.c_int)) for n in I32_FIELDS]


@dataclass
class HostBank:
    arrays: Dict[str, np.ndarray]

    @staticmethod
    def empty(n: int) -> "HostBank":
        a = {k: np.zeros(n, dtype=np.float64) for k in F64_FIELDS}
        a.update({k: np.zeros(n, dtype=np.int32) for k in I32_FIELDS})
        return HostBank(a)

    @staticmethod
    def from_aos(aos: np.ndarray) -> "HostBank":
        assert aos.dtype == PARTICLE_AOS
        return HostBank({k: np.ascontiguousarray(aos[k]) for k in ALL_FIELDS})

    def to_aos(self) -> np.ndarray:
        aos = np.zeros(len(self), dtype=PARTICLE_AOS)
        for k in ALL_FIELDS:
            aos[k] = self.arrays[k]
        return aos

    def __len__(self) -> int:
        return len(self.arrays["x"])

    def __getattr__(self, name):
        try:
            return self.__dict__["arrays"][name]
        except KeyError:
            raise AttributeError(name)

    def copy(self) -> "HostBank":
        return HostBank({k: v.copy() for k, v in self.arrays.items()})

    def slice(self, start: int, count: int) -> "HostBank":
        return HostBank({k: v[start:start + count].copy() for k, v in self.arrays.items()})

    def as_struct(self) -> ParticleSoA:
        """A ParticleSoA whose pointers alias this bank's arrays (keep the bank alive)."""
        s = ParticleSoA()
        for k in F64_FIELDS:
            setattr(s, k, self.arrays[k].ctypes.data_as(C.POINTER(C.c_double)))
        for k in I32_FIELDS:
            setattr(s, k, self.arrays[k].ctypes.data_as(C.POINTER(C.c_int)))
        return s

    def bit_equal(self, other: "HostBank", fields=ALL_FIELDS) -> Dict[str, int]:
        """Number of slots whose bit patterns differ, per field."""
        out = {}
        for k in fields:
            a, b = self.arrays[k], other.arrays[k]
            if a.dtype == np.float64:
                out[k] = int(np.count_nonzero(a.view(np.uint64) != b.view(np.uint64)))
            else:
                out[k] = int(np.count_nonzero(a != b))
        return out
