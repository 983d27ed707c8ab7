#!/usr/bin/env python3
"""Host<->device copy rates of the box (pinned memory), alone and both directions at once."""
import time
import torch

n = 128 * 2**20 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 8 / 1e9
print(f"H2D  128 MiB pinned: {gb / timed(h2d):6.1f} GB/s")
print(f"D2H  128 MiB pinned: {gb / timed(d2h):6.1f} GB/s")
t = timed(both)
print(f"both directions at once: {2 * gb / t:6.1f} GB/s aggregate ({gb / t:.1f} each)")
small = torch.empty(10**6, dtype=torch.float64).pin_memory()
d_small = torch.empty(10**6, dtype=torch.float64, device="cuda")
print(f"H2D  8 MB pinned: {8e-3 / timed(lambda: d_small.copy_(small, non_blocking=True)):6.1f} GB/s")
