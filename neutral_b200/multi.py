"""Particle-sharded timesteps: the host logic of the multi-GPU path.

Histories are independent and their random streams are keyed by the GLOBAL particle index
(``omp3/neutral.c:632-641``), so rank ``r`` of ``R`` transports the contiguous particle range
of :func:`neutral_b200.decks.shard_range` (the split the reference uses for its OpenMP threads,
``omp3/neutral.c:64-74``) against replicated mesh and cross-section tables. The only shared
output is the additive energy-deposition tally, which is CUMULATIVE across timesteps
(``main.c`` never clears it), so what is combined is each timestep's DELTA:

    delta_r = 0;  solve_transport_2d(..., tally = delta_r)           # every rank, its shard
    all_reduce(delta, SUM)                                           # one collective per step
    tally  += delta                                                  # every rank

The all-reduce of timestep ``t`` overlaps the transport of timestep ``t+1`` (two delta buffers):
it is launched asynchronously and only waited for when its buffer is needed again.

The loop is written against a small engine protocol so that the same code drives the CUDA
path (``bench.py``: :class:`GpuShardEngine`, NCCL) and, in the CPU tests, the oracle over
``gloo`` (``tests/test_multi_gloo.py``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Protocol


@dataclass
class StepCounts:
    facets: int
    collisions: int
    processed: int
    census: int = 0
    deaths: int = 0

    @property
    def events(self) -> int:
        return self.facets + self.collisions + self.census


class ShardEngine(Protocol):
    """What one rank must provide. ``k`` selects one of two delta buffers."""

    def step_into_delta(self, tt: int, k: int):
        """One ``solve_transport_2d`` (master_key = tt) of this rank's shard, depositing into
        delta buffer ``k`` (which is zero on entry). Returns the step's counts."""

    def delta_tensor(self, k: int):
        """The torch tensor behind delta buffer ``k`` (what ``all_reduce`` operates on)."""

    def accumulate_and_clear(self, k: int) -> None:
        """tally += delta[k]; delta[k] = 0."""


def run_timesteps(engine: ShardEngine, iterations: int, world: int, dist=None,
                  overlap: bool = True, first_tt: int = 1) -> List:
    """Runs ``iterations`` timesteps of a sharded problem; returns the per-step counts of THIS
    rank. ``dist`` is ``torch.distributed`` (initialised) when ``world > 1``."""
    out = []
    pending = None  # (work handle, buffer index) of the all-reduce still in flight
    for i in range(iterations):
        tt = first_tt + i
        k = i & 1
        out.append(engine.step_into_delta(tt, k))
        if pending is not None:  # timestep tt-1's reduction overlapped this transport
            work, kp = pending
            if work is not None:
                work.wait()
            engine.accumulate_and_clear(kp)
            pending = None
        if world > 1:
            work = dist.all_reduce(engine.delta_tensor(k), async_op=True)
            if overlap and i + 1 < iterations:
                pending = (work, k)
                continue
            work.wait()
        engine.accumulate_and_clear(k)
    return out


def global_counts(local: List, world: int, dist=None, device=None):
    """Sums per-step (facets, collisions, processed, census, deaths) over the ranks."""
    import torch
    t = torch.tensor([[c.facets, c.collisions, c.processed, c.census, c.deaths] for c in local],
                     dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t)
    return [StepCounts(*[int(v) for v in row]) for row in t.tolist()]


class GpuShardEngine:
    """The CUDA path: a :class:`neutral_b200.host.Simulation` plus two device delta buffers."""

    def __init__(self, sim, ncells: int):
        import torch
        self.sim = sim
        self.lib = sim.lib
        self.ncells = ncells
        self.delta = [torch.zeros(ncells, dtype=torch.float64, device="cuda") for _ in range(2)]

    def step_into_delta(self, tt: int, k: int):
        return self.sim.step(tt, tally_ptr=self.delta[k].data_ptr())

    def delta_tensor(self, k: int):
        return self.delta[k]

    def accumulate_and_clear(self, k: int) -> None:
        rc = self.lib.nb200_accumulate_clear(self.sim.tally.ptr, self.delta[k].data_ptr(),
                                             self.ncells)
        if rc != 0:
            raise RuntimeError(self.lib.nb200_last_error().decode())
