"""The multi-GPU host logic on CPU: world size 2 (and 3) over ``gloo``.

Each rank drives ``neutral_b200.multi.run_timesteps`` - the loop ``bench.py`` uses under
torchrun - with the oracle port as its per-rank engine (the CUDA engine needs a GPU): the
contiguous particle shard of ``shard_range``, global RNG keys, per-timestep tally deltas
combined by one all-reduce, with and without overlapping the reduction with the next
timestep. The combined result must equal the single-rank run: exact event counts per
timestep, bit-identical shard banks, tally within 1e-12 (summation order only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from neutral_b200.decks import build_problem, shard_range  # noqa: E402
from neutral_b200.multi import StepCounts, global_counts, run_timesteps  # noqa: E402


class OracleShardEngine:
    """ShardEngine over the CPU oracle (test infrastructure)."""

    def __init__(self, prob, rank, world):
        from oracle.oracle import OraclePort
        self.port = OraclePort()
        self.prob = prob
        d = prob.deck
        self.pid0, self.count = shard_range(d.nparticles, rank, world)
        self.bank = self.port.inject(prob, self.pid0, self.count)
        self.tally = np.zeros(d.nx * d.ny)
        self.delta = [torch.zeros(d.nx * d.ny, dtype=torch.float64)
                      for _ in range(getattr(self, "nbuffers", 2))]

    def step_into_delta(self, tt, k):
        assert not self.delta[k].any(), "delta buffer must be clear on entry"
        before = int(np.count_nonzero(self.bank.dead == 0))
        f, c, p = self.port.step(self.prob, self.bank, tt, self.delta[k].numpy(), pid0=self.pid0)
        after = int(np.count_nonzero(self.bank.dead == 0))
        assert p == before
        return StepCounts(f, c, p, census=after, deaths=before - after)

    def delta_tensor(self, k):
        return self.delta[k]

    def accumulate_and_clear(self, k):
        self.tally += self.delta[k].numpy()
        self.delta[k].zero_()


class AsyncOracleShardEngine(OracleShardEngine):
    """The protocol of the CUDA engine (three buffers, folds queued behind the collective and
    taken off the loop's critical path), with the queue drained lazily on the host."""

    nbuffers = 3

    def __init__(self, prob, rank, world):
        super().__init__(prob, rank, world)
        self.queue = []
        self.max_queued = 0

    def step_begin(self, tt, k):
        self._begun = (tt, k)

    def step_end(self):
        # the collective of this step was launched before its transport "finished": run the
        # transport now and make the queued fold wait for a reduction of the finished delta
        tt, k = self._begun
        counts = self.step_into_delta(tt, k)
        kq, _ = self.queue.pop()
        assert kq == k
        self.queue.append((k, dist.all_reduce(self.delta[k], async_op=True)))
        return counts

    def fold_async(self, k, work):
        if work is not None and hasattr(self, "_begun") and self._begun[1] == k:
            work.wait()  # a reduction of the still-empty buffer: harmless, discarded
        self.queue.append((k, work))
        self.max_queued = max(self.max_queued, len(self.queue))

    def _run(self, n):
        for k, work in self.queue[:n]:
            work.wait()
            self.accumulate_and_clear(k)
        del self.queue[:n]

    def acquire(self, k):
        ks = [q[0] for q in self.queue]
        if k in ks:
            self._run(ks.index(k) + 1)

    def drain(self):
        self._run(len(self.queue))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, deck, overlap, out, async_fold=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prob = build_problem(deck)
        eng = (AsyncOracleShardEngine if async_fold else OracleShardEngine)(prob, rank, world)
        local = run_timesteps(eng, prob.deck.iterations, world, dist, overlap=overlap)
        if async_fold:
            assert not eng.queue and eng.max_queued >= min(2, prob.deck.iterations)
        total = global_counts(local, world, dist)
        # every rank must hold the same cumulative tally
        t = torch.from_numpy(eng.tally.copy())
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on the reduced tally"
        np.savez(os.path.join(out, f"rank{rank}.npz"), tally=eng.tally,
                 counts=np.array([[c.facets, c.collisions, c.processed, c.census, c.deaths]
                                  for c in total]),
                 pid0=eng.pid0, **{f"bank_{k}": v for k, v in eng.bank.arrays.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,deck,overlap,async_fold",
                         [(2, "mixed_small", True, False), (2, "csp_small", False, False),
                          (3, "split_small", True, False), (2, "csp_small", True, True)])
def test_sharded_timesteps_over_gloo(tmp_path, port, world, deck, overlap, async_fold):
    mp.spawn(_worker, args=(world, _free_port(), deck, overlap, str(tmp_path), async_fold),
             nprocs=world, join=True)
    prob = build_problem(deck)
    d = prob.deck
    bank = port.inject(prob)
    tally = np.zeros(d.nx * d.ny)
    want = []
    for tt in range(1, d.iterations + 1):
        before = int(np.count_nonzero(bank.dead == 0))
        f, c, p = port.step(prob, bank, tt, tally)
        after = int(np.count_nonzero(bank.dead == 0))
        want.append([f, c, p, after, before - after])
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert np.array_equal(got["counts"], np.array(want)), f"rank {r}: global counts"
        pid0, count = shard_range(d.nparticles, r, world)
        assert int(got["pid0"]) == pid0
        for k, v in bank.arrays.items():
            assert got[f"bank_{k}"].tobytes() == v[pid0:pid0 + count].tobytes(), (r, k)
        t = got["tally"]
        assert np.all(np.abs(t - tally) <= 1e-12 * np.maximum(np.abs(t), np.abs(tally)))


def test_shard_ranges_tile_the_bank():
    for n, world in [(10, 3), (1_000_000, 8), (7, 8), (100_000_000, 8)]:
        nxt = 0
        for r in range(world):
            first, count = shard_range(n, r, world)
            assert first == nxt and count in (n // world, n // world + 1)
            nxt = first + count
        assert nxt == n


# ------------------------------------------------------------------------------------------
# The owned-slice protocol of the library's own collective (csrc/nb_group.cuh), over gloo
# ------------------------------------------------------------------------------------------

def _owned_worker(rank, world, port_no, deck, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from neutral_b200.multi import OwnedSliceTally
        prob = build_problem(deck)
        d = prob.deck
        eng = OracleShardEngine(prob, rank, world)
        group = OwnedSliceTally(d.nx * d.ny, rank, world, dist)
        tally = torch.zeros(d.nx * d.ny, dtype=torch.float64)
        counts = []
        for tt in range(1, d.iterations + 1):
            counts.append(eng.step_into_delta(tt, 0))
            group.reduce_fold(eng.delta[0])       # every timestep: reduce-scatter + fold
            if tt == 2:
                group.flush(tally)                # somebody looks at the tally mid-run
        group.flush(tally)
        assert not bool(group.owned.any())
        g = global_counts(counts, world, dist)
        if rank == 0:
            out.put(([(c.facets, c.collisions, c.processed) for c in g], tally.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,deck", [(2, "mixed_small"), (3, "csp_small")])
def test_owned_slice_protocol_equals_single_rank(world, deck):
    """Reduce-scatter into owned slices every timestep + all-gather when the tally is looked
    at (the protocol of csrc/nb_group.cuh, ragged last slice included: 80 x 80 = 6400 cells
    over 3 ranks -> chunk 2134) reproduces the single-rank run: exact counts, tally to 1e-12."""
    from neutral_b200.multi import owned_slice_layout
    from oracle.oracle import OraclePort
    assert owned_slice_layout(16_000_000, 8) == (2_000_000, 16_000_000)
    assert owned_slice_layout(6400, 3) == (2134, 6402)
    assert owned_slice_layout(9, 2) == (6, 12)
    prob = build_problem(deck)
    d = prob.deck
    port = OraclePort()
    bank = port.inject(prob)
    want_tally = np.zeros(d.nx * d.ny)
    want = [port.step(prob, bank, tt, want_tally) for tt in range(1, d.iterations + 1)]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_owned_worker, args=(r, world, port_no, deck, out))
             for r in range(world)]
    for p in procs:
        p.start()
    got_counts, got_tally = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got_counts == want
    assert np.all(np.abs(got_tally - want_tally) <=
                  1e-12 * np.maximum(np.abs(got_tally), np.abs(want_tally)))
