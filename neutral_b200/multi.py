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

The all-reduce of timestep ``t`` overlaps the transport of timestep ``t+1``: it is launched
asynchronously and its buffer is only folded into the tally when the buffer is needed again.
An engine with three delta buffers and a side stream (:class:`GpuShardEngine`) also takes the
fold itself off the critical path: ``tally += delta; delta = 0`` is queued behind the
all-reduce on the side stream and runs beside the next transport, which deposits into the
third buffer.

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
    """What one rank must provide. ``k`` selects one of ``nbuffers`` delta buffers (attribute,
    default 2). An engine that can fold asynchronously also provides ``fold_async(k, work)``
    (queue "wait for ``work``, tally += delta[k], delta[k] = 0" without blocking the host;
    with ``step_begin(tt, k)`` / ``step_end()`` the timestep itself is only enqueued first, so
    the collective is launched while the transport runs), ``acquire(k)`` (order the next use of buffer ``k`` behind its queued fold) and ``drain()``
    (order everything that follows behind all queued folds)."""

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
    nbuf = getattr(engine, "nbuffers", 2)
    async_fold = overlap and world > 1 and hasattr(engine, "fold_async")
    pending = []  # (step index, buffer, work handle): reductions in flight, fold still owed

    def fold_through(last_index: int) -> None:
        while pending and pending[0][0] <= last_index:
            _, kp, work = pending.pop(0)
            if work is not None:
                work.wait()
            engine.accumulate_and_clear(kp)

    for i in range(iterations):
        k = i % nbuf
        fold_through(i - nbuf)  # buffer k is about to be deposited into again
        if async_fold:
            engine.acquire(k)
        if async_fold and hasattr(engine, "step_begin"):
            # the collective and the fold are queued behind the transport while it runs, so
            # their launch latency is off the critical path too
            engine.step_begin(first_tt + i, k)
            engine.fold_async(k, dist.all_reduce(engine.delta_tensor(k), async_op=True))
            out.append(engine.step_end())
            continue
        out.append(engine.step_into_delta(first_tt + i, k))
        work = dist.all_reduce(engine.delta_tensor(k), async_op=True) if world > 1 else None
        if async_fold:
            engine.fold_async(k, work)
        elif overlap and world > 1:
            fold_through(i - 1)  # the previous reduction overlapped this transport
            pending.append((i, k, work))
        else:
            pending.append((i, k, work))
            fold_through(i)
    fold_through(iterations)
    if async_fold:
        engine.drain()
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
    """The CUDA path: a :class:`neutral_b200.host.Simulation`, three device delta buffers and a
    side stream on which reduced deltas are folded into the tally."""

    nbuffers = 3

    def __init__(self, sim, ncells: int):
        import torch
        self.torch = torch
        self.sim = sim
        self.lib = sim.lib
        self.ncells = ncells
        self.delta = [torch.zeros(ncells, dtype=torch.float64, device="cuda")
                      for _ in range(self.nbuffers)]
        self.side = torch.cuda.Stream()
        self.folded = [torch.cuda.Event() for _ in range(self.nbuffers)]

    def step_into_delta(self, tt: int, k: int):
        return self.sim.step(tt, tally_ptr=self.delta[k].data_ptr())

    def step_begin(self, tt: int, k: int) -> None:
        """Enqueues the timestep and returns (``Simulation.step(defer=True)``)."""
        self.sim.step(tt, tally_ptr=self.delta[k].data_ptr(), defer=True)

    def step_end(self):
        return self.sim.step_finish()

    def delta_tensor(self, k: int):
        return self.delta[k]

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(self.lib.nb200_last_error().decode())

    def accumulate_and_clear(self, k: int) -> None:
        self._check(self.lib.nb200_accumulate_clear(self.sim.tally.ptr, self.delta[k].data_ptr(),
                                                    self.ncells))

    def fold_async(self, k: int, work) -> None:
        torch = self.torch
        if work is None:  # no collective in front: order the fold behind the transport itself
            self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            if work is not None:
                work.wait()  # the side stream waits for the collective, the host does not
            self._check(self.lib.nb200_accumulate_clear_async(
                self.sim.tally.ptr, self.delta[k].data_ptr(), self.ncells,
                self.side.cuda_stream))
            self.folded[k].record(self.side)

    def acquire(self, k: int) -> None:
        self.torch.cuda.current_stream().wait_event(self.folded[k])

    def drain(self) -> None:
        self.torch.cuda.current_stream().wait_stream(self.side)


# ----------------------------------------------------------------------------------------
# The owned-slice protocol of the library's own collective (csrc/nb_group.cuh), restated over
# torch.distributed. The CUDA path does this inside libneutral_b200.so with one peer-memory
# kernel per rank and timestep; this restatement is what the CPU tests drive over `gloo` (with
# the oracle as the per-rank engine) to pin the slicing, the padding and the flush.
# ----------------------------------------------------------------------------------------

def owned_slice_layout(ncells: int, nranks: int):
    """(chunk, padded): cells per owned slice - even, so that 16-byte vector accesses never
    straddle two owners - and the padded tally length nranks * chunk (group_layout, capi.cu)."""
    chunk = (ncells + nranks - 1) // nranks
    chunk = (chunk + 1) & ~1
    return chunk, chunk * nranks


class OwnedSliceTally:
    """Rank ``rank``'s part of a tally shared by ``world`` ranks: it owns cells
    [rank * chunk, (rank + 1) * chunk) of the cumulative tally. ``reduce_fold`` takes every
    rank's delta of one timestep (reduce-scatter, summed in rank order, folded into the owned
    slice); ``flush`` adds all owned slices into a caller-visible tally (all-gather) and clears
    them."""

    def __init__(self, ncells: int, rank: int, world: int, dist=None):
        import torch
        self.torch = torch
        self.ncells, self.rank, self.world, self.dist = ncells, rank, world, dist
        self.chunk, self.padded = owned_slice_layout(ncells, world)
        self.owned = torch.zeros(self.chunk, dtype=torch.float64)

    def reduce_fold(self, delta) -> None:
        """delta: this rank's per-timestep tally (ncells doubles, torch); cleared afterwards."""
        torch = self.torch
        padded = torch.zeros(self.padded, dtype=torch.float64)
        padded[:self.ncells] = delta
        if self.world > 1:
            # reduce-scatter spelled with point-to-point reductions (gloo has no reduce_scatter):
            # slice r of every rank's delta is summed on rank r
            for r in range(self.world):
                part = padded[r * self.chunk:(r + 1) * self.chunk].clone()
                self.dist.reduce(part, dst=r)
                if r == self.rank:
                    self.owned += part
        else:
            self.owned += padded[:self.chunk]
        delta.zero_()

    def flush(self, tally) -> None:
        """tally (ncells doubles, torch) += every rank's owned slice; owned slices cleared."""
        torch = self.torch
        if self.world > 1:
            parts = [torch.zeros(self.chunk, dtype=torch.float64) for _ in range(self.world)]
            self.dist.all_gather(parts, self.owned)
            whole = torch.cat(parts)
        else:
            whole = self.owned
        tally += whole[:self.ncells]
        self.owned.zero_()
