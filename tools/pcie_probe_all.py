#!/usr/bin/env python3
"""Host<->device copy rate of the WHOLE box: every rank (one per GPU, torchrun) copies pinned
buffers up and down at the same time, as bench.py's end-to-end loop does. Says what the ranks
get each and in aggregate - the ceiling of any end-to-end figure that moves B bytes per step:
N * B / aggregate.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29519 tools/pcie_probe_all.py
"""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cores = sorted(os.sched_getaffinity(0))
per = len(cores) // world
if per >= 1:
    os.sched_setaffinity(0, cores[local * per:(local + 1) * per])

n = 200 * 10**6 // 8  # ~ the bytes bench.py moves each way per csp deck run
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.empty(n, dtype=torch.float64, device="cuda")
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_up, do_down, reps=8):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if do_up:
            with torch.cuda.stream(up):
                d_a.copy_(h_in, non_blocking=True)
        if do_down:
            with torch.cuda.stream(down):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) / reps


gb = n * 8 / 1e9
for name, u, d in (("H2D only", True, False), ("D2H only", False, True), ("both ways", True, True)):
    run(u, d, 2)
    s = run(u, d)
    moved = gb * (int(u) + int(d))
    if rank == 0:
        print(f"{world} ranks, {name:9s}: {moved / s:6.1f} GB/s per rank, {world * moved / s:7.1f} GB/s "
              f"aggregate ({s * 1e3:.2f} ms per {moved:.2f} GB)", flush=True)
dist.destroy_process_group()
