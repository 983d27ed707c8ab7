#!/usr/bin/env python3
"""What combining the tally costs a particle-sharded run (torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/group_breakdown.py [deck]

One deck run per mode, device time (max over ranks), best of 3, next to rank 0's own history-
and sort-kernel time: (a) transport only - no collective at all, every rank into its own tally;
(b) the library group with its peer-memory reduce-scatter kernel, reduced on demand (the default:
once per deck run, when the tally is read) and every timestep (beside the next timestep's
transport), each followed by the all-gather into the caller's tally; (c) the same with NCCL's
reduce-scatter inside the library; (d) round 1's host loop: torch all-reduce of the whole delta
+ fold on a side stream (neutral_b200/multi.py).
"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import connect_group  # noqa: E402
from neutral_b200.decks import build_problem, load_deck  # noqa: E402
from neutral_b200.host import Simulation, _check, _soa_p, load_library  # noqa: E402
from neutral_b200.multi import GpuShardEngine, run_timesteps  # noqa: E402

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
snap = _soa_p()
injected = sim.bank_to_host()  # must outlive as_struct(): the struct aliases its arrays
st = injected.as_struct()
_check(lib.nb200_bank_create(C.byref(st), sim.count, sim.pid0, C.byref(snap)), "snapshot")
rows = []


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        rows.clear()
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


def report(name, ms):
    k = torch.tensor([sum(r.kernel_ns for r in rows) / 1e6, sum(r.sort_ns for r in rows) / 1e6],
                     device="cuda", dtype=torch.float64)
    dist.all_reduce(k, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{deck.name} x{world}: {name:44s} {ms:8.3f} ms per deck run   slowest rank: history "
              f"{float(k[0]):7.3f} ms, sort phase {float(k[1]):6.3f} ms", flush=True)


def pipelined():
    rows.extend(sim.run_pipelined())


def pipelined_and_sync():
    rows.extend(sim.run_pipelined())
    sim.tally_sync()


report("transport only (no collective)", timed(pipelined))
for flavour, every, name in ((1, 0, "library group: peer-memory kernel, reduce on demand"),
                             (1, 1, "library group: peer-memory kernel, every timestep"),
                             (0, 0, "library group: NCCL, reduce on demand"),
                             (0, 1, "library group: NCCL, every timestep")):
    lib.nb200_set_option(b"collective", flavour)
    lib.nb200_set_option(b"tally_reduce_every", every)
    what = connect_group(lib, dist, torch, world, rank, ncells)
    report(f"{name}", timed(pipelined_and_sync))
    lib.nb200_mp_finalize()
lib.nb200_set_option(b"collective", 1)
lib.nb200_set_option(b"tally_reduce_every", 0)
engine = GpuShardEngine(sim, ncells)
report("round-1 host loop: torch all-reduce + fold",
       timed(lambda: rows.extend(run_timesteps(engine, d.iterations, world, dist, overlap=True))))
dist.destroy_process_group()
