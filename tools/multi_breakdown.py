#!/usr/bin/env python3
"""Where the multi-GPU timestep goes (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_breakdown.py [deck]

Times one deck run per mode (device time, max over ranks): transport only (no collective),
all-reduce after every timestep without overlap, the overlapped loop of neutral_b200/multi.py,
and a lone all-reduce / accumulate of the tally-sized buffer.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neutral_b200.decks import build_problem, load_deck  # noqa: E402
from neutral_b200.host import Simulation, _check, _soa_p, load_library  # noqa: E402
from neutral_b200.multi import GpuShardEngine, run_timesteps  # noqa: E402
import ctypes as C  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_library(build=False)
lib.nb200_set_option(b"print", 0)
deck = load_deck(sys.argv[1] if len(sys.argv) > 1 else "csp")
prob = build_problem(deck, nparticles=deck.nparticles * world)
d = prob.deck
ncells = d.nx * d.ny
sim = Simulation(prob, rank=rank, nranks=world, per_particle_counters=False)
sim.inject()
start = sim.bank_to_host()
snap = _soa_p()
st = start.as_struct()
_check(lib.nb200_bank_create(C.byref(st), sim.count, sim.pid0, C.byref(snap)), "snapshot")
engine = GpuShardEngine(sim, ncells)


last_rows = []


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        last_rows.clear()
        _check(lib.nb200_bank_copy(sim.bank, snap), "copy")
        sim.tally.zero()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t[0]) if best is None else min(best, float(t[0]))
    return best


def transport_only():
    for tt in range(1, d.iterations + 1):
        last_rows.append(engine.step_into_delta(tt, 0))


def no_overlap():
    last_rows.extend(run_timesteps(engine, d.iterations, world, dist, overlap=False))


def overlapped():
    last_rows.extend(run_timesteps(engine, d.iterations, world, dist, overlap=True))


dummy = torch.zeros(ncells, dtype=torch.float64, device="cuda")


def transport_beside_allreduce():
    """Transport with an unrelated all-reduce of the same size in flight: no fold, no
    dependency - isolates what sharing the GPU with the collective costs the kernels."""
    works = []
    for tt in range(1, d.iterations + 1):
        last_rows.append(engine.step_into_delta(tt, 0))
        works.append(dist.all_reduce(dummy, async_op=True))
    for w in works:
        w.wait()


def transport_beside_fold():
    for tt in range(1, d.iterations + 1):
        last_rows.append(engine.step_into_delta(tt, 0))
        engine.fold_async(1, None)
    engine.drain()


buf = torch.zeros(ncells, dtype=torch.float64, device="cuda")


def lone_allreduce():
    for _ in range(d.iterations):
        dist.all_reduce(buf)


def lone_accumulate():
    for _ in range(d.iterations):
        engine.accumulate_and_clear(0)


rows = [("transport only", transport_only), ("all-reduce, no overlap", no_overlap),
        ("all-reduce overlapped", overlapped),
        ("transport beside all-reduce", transport_beside_allreduce),
        ("transport beside fold", transport_beside_fold), (f"{d.iterations} lone all-reduces", lone_allreduce),
        (f"{d.iterations} lone accumulate+clear", lone_accumulate)]
for name, fn in rows:
    ms = timed(fn)
    if rank == 0:
        k_ms = sum(r.kernel_ns + r.sort_ns for r in last_rows) / 1e6
        print(f"{deck.name} x{world}: {name:28s} {ms:9.3f} ms   (rank 0 transport kernels "
              f"{k_ms:8.3f} ms)", flush=True)
dist.destroy_process_group()
