#!/usr/bin/env python3
"""bench.py - particle events/s of neutral's particle-history hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--deck csp] [--particles P]

One "step" = one complete run of the deck named in ``config.workload`` (default csp: 4000 x
4000 mesh, 1e6 particles per GPU, 10 timesteps of solve_transport_2d) from the freshly
injected bank. Prints ONE JSON line (rank 0):

* ``value``     whole-job events/s ( facets + collisions + census over all ranks and the K
                timed steps / device time, max over ranks ), inputs resident in HBM;
* ``e2e``       the same metric with HOST buffers: every step uploads the deck (mesh,
                cross sections, bank) from pinned host memory, runs the timesteps through
                solve_transport_2d, and reads tally and bank back (double-buffered: the
                copies of neighbouring steps overlap a step's transport);
* ``roofline``  the history kernel against the measured HBM peak, algorithmic bytes per
                event from SURVEY.md 8d (facet 200 B, collision 176 B, census 192 B, fatal
                collision +16 B);
* ``cpu_baseline`` the unmodified reference omp3 build (oracle/_ref) on this box's host
                cores, on a bounded sample of the same deck (N=1 only).

``--impl reference`` times the reference's own CPU implementation instead (rank 0 only).
N > 1 (torchrun): weak scaling - every rank transports ``deck.nparticles`` particles of an
N-times larger global bank (global RNG keys), and the per-timestep tally deltas are
all-reduced with NCCL.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle_events_per_sec"
UNIT = "events/s"
# algorithmic bytes per event of the event-based model
B_FACET, B_COLLISION, B_CENSUS, B_DEATH_EXTRA = 200.0, 176.0, 192.0, 16.0
RED_PEAK_PER_S = 1.97e11  # measured best case: red.global.add.f64, spread addresses, L2-resident
                          # footprint (tools/microbench/red_rate.cu, profiles/r01/red_rate_v2.txt)
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--deck", default="csp")
    ap.add_argument("--particles", type=int, default=0, help="override particles per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opts", default="", help="library options, e.g. pipeline=0,fast_div=0")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks --

class ClockSampler:
    """Samples nvidia-smi every 100 ms (B200_PROFILING.md clocks line). It is started before
    the warm-up so that it is already reporting when the timed regions begin; only samples
    that arrived inside a timed region (`window(t0, t1)`) count."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []  # (arrival time, csv line)
        self.windows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # a sample describes the 100 ms before it arrived
        inside = [ln for t, ln in self.lines
                  if any(t0 <= t <= t1 + 0.1 for t0, t1 in self.windows)]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------- reference (CPU) arm --

def reference_rate(deck_name: str, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times the unmodified reference omp3 build (oracle/_ref, else the oracle port) on this
    host, all cores, on a bounded sample of the deck. Returns a dict for the JSON line."""
    from neutral_b200.decks import build_problem, load_deck
    from oracle.oracle import OraclePort, ReferenceOmp3

    cores = len(os.sched_getaffinity(0))
    # All the host threads the box has, whatever the launcher exported: torchrun sets
    # OMP_NUM_THREADS=1 for its workers, which would time the reference on one core while
    # reporting sixteen. NB200_REF_THREADS overrides.
    cores = int(os.environ.get("NB200_REF_THREADS", cores))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    kind = "reference" if ReferenceOmp3.available() else "port"
    eng = ReferenceOmp3() if kind == "reference" else OraclePort()
    try:  # the OpenMP runtime may have read the environment before this point (torch loads one)
        C.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    deck = load_deck(deck_name)

    def run_once(nparticles):
        prob = build_problem(deck, nparticles=nparticles)
        d = prob.deck
        tally = np.zeros(d.nx * d.ny)
        bank = eng.inject(prob)
        dead = (lambda: bank["dead"]) if kind == "reference" else (lambda: bank.dead)
        events, seconds = 0, 0.0
        for tt in range(1, d.iterations + 1):
            t0 = time.perf_counter()
            counts = eng.step(prob, bank, tt, tally)  # (facets, collisions[, processed])
            seconds += time.perf_counter() - t0
            # census events = particles still alive after the step
            events += counts[0] + counts[1] + int(np.count_nonzero(dead() == 0))
        return events, seconds

    # calibrate on a small sample, then size the real one for ~target_seconds per step
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints "Particles N" every timestep
    try:
        n0 = min(deck.nparticles, 2000 * cores)
        ev0, s0 = run_once(n0)
        rate0 = ev0 / max(s0, 1e-9)
        per_particle = ev0 / n0
        n = int(min(deck.nparticles, max(n0, rate0 * target_seconds / per_particle)))
        times, events = [], 0
        for i in range(warmup + steps):
            ev, s = run_once(n)
            if i >= warmup:
                times.append(s)
                events += ev
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
    total = sum(times)
    return {"value": events / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{deck.name}: {n} of {deck.nparticles} particles, {deck.iterations} "
                      f"timesteps, {deck.nx}x{deck.ny} mesh, {steps} run(s) of "
                      f"{total / max(steps, 1):.1f} s, OMP threads={cores}",
            "ms_per_step": 1e3 * total / max(steps, 1), "events_per_step": events / max(steps, 1)}


def run_reference_arm(args, rank: int):
    if rank != 0:
        return
    budget = 150.0 / max(args.steps + args.warmup, 1)
    res = reference_rate(args.deck, target_seconds=min(10.0, budget), steps=args.steps,
                         warmup=args.warmup)
    from neutral_b200.decks import load_deck
    deck = load_deck(args.deck)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{deck.name}.params (bounded sample, see cpu_baseline.sample)",
                   "mesh": [deck.nx, deck.ny], "timesteps_per_step": deck.iterations},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- b200 (GPU) arm --

def algorithmic_bytes(results):
    return sum(r.facets * B_FACET + r.collisions * B_COLLISION + r.census * B_CENSUS +
               r.deaths * B_DEATH_EXTRA for r in results)


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(deck_name):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(deck_name, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def fused_history_bytes(results):
    """Lower bound on the bytes a fused-history kernel must move (B_hist): the
    80-byte record in and out once per particle-timestep, the entered cell's density and the
    tally read-modify-write per facet, the tally RMW per census or death."""
    return sum(r.processed * 160.0 + r.facets * 24.0 + (r.census + r.deaths) * 16.0
               for r in results)


def run_b200_arm(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist

    from neutral_b200.bank import HostBank
    from neutral_b200.decks import build_problem, load_deck
    from neutral_b200.host import Simulation, _check, _soa_p, load_library
    from neutral_b200.multi import GpuShardEngine, run_timesteps

    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries the one JSON line: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = load_library(build=False)
    lib.initialise_devices.restype = None
    lib.nb200_set_option(b"print", 0)
    opts = dict(kv.split("=") for kv in args.opts.split(",") if kv)
    for k, v in opts.items():
        if lib.nb200_set_option(k.encode(), int(v)) < -1:
            sys.exit(f"unknown library option {k}")
    pipeline = int(opts.get("pipeline", 1))

    deck = load_deck(args.deck)
    per_gpu = args.particles or deck.nparticles
    prob = build_problem(deck, nparticles=per_gpu * world)  # weak scaling: global bank grows
    d = prob.deck
    ncells = d.nx * d.ny

    sim = Simulation(prob, rank=rank, nranks=world, per_particle_counters=False)
    sim.inject()
    start_bank = sim.bank_to_host()  # host copy of the freshly injected shard
    # resident snapshot the timed steps restart from
    snap = _soa_p()
    st = start_bank.as_struct()
    _check(lib.nb200_bank_create(C.byref(st), sim.count, sim.pid0, C.byref(snap)), "snapshot")

    # N > 1: per-timestep tally deltas, one NCCL all-reduce each, overlapped with the next
    # timestep's transport (neutral_b200/multi.py)
    engine = GpuShardEngine(sim, ncells) if world > 1 else None

    def timesteps():
        if world > 1:
            return run_timesteps(engine, d.iterations, world, dist)
        return [sim.step(tt) for tt in range(1, d.iterations + 1)]

    def one_step():
        """One deck run from the injected state; returns the list of StepResults."""
        _check(lib.nb200_bank_copy(sim.bank, snap), "bank_copy")
        sim.tally.zero()
        return timesteps()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()

    launches0 = lib.nb200_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    wall0 = time.time()
    ev0.record()
    timed = []
    for _ in range(args.steps):
        timed += one_step()
    ev1.record()
    fence()
    sampler.window(wall0, time.time())
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = lib.nb200_kernel_launches() - launches0

    events = sum(r.events for r in timed)
    kernel_ns = sum(r.kernel_ns for r in timed)
    sort_ns = sum(r.sort_ns for r in timed)
    hist_launches = len(timed)
    alg_bytes = algorithmic_bytes(timed)
    tally_sum = float(torch.from_numpy(sim.tally_to_host()).sum()) if rank == 0 else 0.0

    # ---- e2e: host buffers in, host buffers out, every step ---------------------------
    # Every step uploads ALL of its inputs (mesh, edges, cross-section tables, bank) from
    # pinned host memory and downloads its results (tally, bank) to pinned host memory. The
    # steps are independent, so the transfers are double-buffered the way a production host
    # would: two device-side working sets; while step i transports on set i%2, the upload of
    # step i+1 and the download of step i-1 run on their own copy streams. Nothing is skipped:
    # the timed region holds K full uploads, K full runs and K full downloads.
    e2e = None
    if not args.no_e2e:
        from neutral_b200.bank import ALL_FIELDS, ParticleSoA
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_density, h_ex, h_ey = pin(prob.density.ravel()), pin(prob.edgex), pin(prob.edgey)
        h_cs = [pin(a) for pair in (prob.cs_scatter, prob.cs_absorb) for a in pair]
        h_bank = {k: pin(v) for k, v in start_bank.arrays.items()}
        nbytes = lambda t: t.numel() * t.element_size()

        class WorkingSet:
            def __init__(self, s):
                self.sim = s
                self.engine = GpuShardEngine(s, ncells) if world > 1 else None
                self.view = ParticleSoA()
                _check(lib.nb200_bank_view(s.bank, C.byref(self.view)), "bank_view")
                self.inputs = [(s.density, h_density), (s.edgex, h_ex), (s.edgey, h_ey)] + \
                    list(zip(s._cs_arrays, h_cs))
                self.out_tally = torch.empty(ncells, dtype=torch.float64).pin_memory()
                self.out_bank = {k: torch.empty_like(v).pin_memory() for k, v in h_bank.items()}
                self.uploaded = torch.cuda.Event()
                self.downloaded = torch.cuda.Event()

            def field_ptr(self, k):
                return C.cast(getattr(self.view, k), C.c_void_p).value

        sim_b = Simulation(prob, rank=rank, nranks=world, per_particle_counters=False)
        sim_b.load_bank(start_bank)
        sets = [WorkingSet(sim), WorkingSet(sim_b)]
        up, down = torch.cuda.Stream(), torch.cuda.Stream()
        h2d = sum(nbytes(t) for _, t in sets[0].inputs) + sum(nbytes(t) for t in h_bank.values())
        d2h = ncells * 8 + sum(nbytes(t) for t in h_bank.values())

        def enqueue_upload(ws):
            up.wait_event(ws.downloaded)  # the set's previous results must have left it
            for dev, host in ws.inputs:
                _check(lib.nb200_memcpy_h2d_async(dev.ptr, host.data_ptr(), nbytes(host),
                                                  up.cuda_stream), "h2d")
            for k in ALL_FIELDS:
                _check(lib.nb200_memcpy_h2d_async(ws.field_ptr(k), h_bank[k].data_ptr(),
                                                  nbytes(h_bank[k]), up.cuda_stream), "h2d bank")
            ws.uploaded.record(up)

        trace = {} if os.environ.get("NB200_E2E_TRACE") else None

        def lap(name, t0):
            if trace is not None:
                trace[name] = trace.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        def transport(ws):
            t = time.perf_counter()
            torch.cuda.current_stream().wait_event(ws.uploaded)
            if trace is not None:
                ws.uploaded.synchronize()
                t = lap("wait_upload", t)
            _check(lib.nb200_bank_import(ws.sim.bank), "bank_import")
            ws.sim.tally.zero()
            t = lap("import+zero (enqueue)", t)
            if world > 1:
                out = run_timesteps(ws.engine, d.iterations, world, dist)
            else:
                out = []
                for tt in range(1, d.iterations + 1):
                    out.append(ws.sim.step(tt))
                    t = lap("step %d" % tt if tt <= 2 else "steps 3..", t)
            _check(lib.nb200_bank_export(ws.sim.bank), "bank_export")  # synchronises
            lap("export", t)
            return out

        def enqueue_download(ws):
            _check(lib.nb200_memcpy_d2h_async(ws.out_tally.data_ptr(), ws.sim.tally.ptr,
                                              ncells * 8, down.cuda_stream), "d2h")
            for k in ALL_FIELDS:
                _check(lib.nb200_memcpy_d2h_async(ws.out_bank[k].data_ptr(), ws.field_ptr(k),
                                                  nbytes(h_bank[k]), down.cuda_stream), "d2h bank")
            ws.downloaded.record(down)

        def e2e_run(nsteps):
            out = []
            enqueue_upload(sets[0])
            for i in range(nsteps):
                if i + 1 < nsteps:
                    enqueue_upload(sets[(i + 1) & 1])
                out += transport(sets[i & 1])
                enqueue_download(sets[i & 1])
            up.synchronize()
            down.synchronize()
            return out

        e2e_run(min(args.warmup, 2))
        fence()
        t0 = time.perf_counter()
        e2e_res = e2e_run(args.steps)
        fence()
        e2e_s = time.perf_counter() - t0
        if trace is not None:
            print("e2e host trace (ms per step):", {k: round(1e3 * v / (args.steps + min(args.warmup, 2)), 3)
                                                    for k, v in trace.items()}, file=sys.stderr)
        sampler.window(time.time() - e2e_s, time.time())
        last = sets[(args.steps - 1) & 1]
        e2e = {"events": sum(r.events for r in e2e_res), "seconds": e2e_s, "h2d": h2d,
               "hist_ms": sum(r.kernel_ns for r in e2e_res) / 1e6 / args.steps,
               "sort_ms": sum(r.sort_ns for r in e2e_res) / 1e6 / args.steps,
               "d2h": d2h, "tally_sum": float(last.out_tally.sum()),
               "live_out": int((last.out_bank["dead"] == 0).sum())}

    clocks = sampler.stop() if rank == 0 else None  # samples of both timed regions

    # ---- reduce over ranks -------------------------------------------------------------
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e["seconds"] if e2e else 0.0], device="cuda",
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([events, e2e["events"] if e2e else 0, launches, kernel_ns, alg_bytes],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        elapsed_ms, e2e_seconds = float(t[0]), float(t[1])
        events_all, e2e_events_all, launches_all = float(c[0]), float(c[1]), int(c[2])
    else:
        e2e_seconds = e2e["seconds"] if e2e else 0.0
        events_all, e2e_events_all, launches_all = float(events), \
            float(e2e["events"]) if e2e else 0.0, int(launches)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    # dominant kernel = the history kernel; per-launch figures of THIS rank (rank 0)
    ach_gbs = (alg_bytes / max(kernel_ns, 1))  # bytes per ns == GB/s
    roofline = {
        "bound": "hbm", "kernel": "k_history" if pipeline else "k_history_direct",
        "achieved": ach_gbs, "peak": peak,
        "unit": "GB/s", "frac": ach_gbs / peak, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes / max(hist_launches, 1),
        "avg_launch_ms": kernel_ns / max(hist_launches, 1) / 1e6,
        "launches_timed": hist_launches,
        "kernel_share_of_step": (kernel_ns / 1e6) / max(elapsed_ms, 1e-9),
        "sort_phase_share_of_step": (sort_ns / 1e6) / max(elapsed_ms, 1e-9),
        "traffic": ncu_traffic_per_launch(deck.name),
        # B_evt is the event-stepped model's traffic (the contract figure); this
        # design keeps a history in registers for a whole timestep, so `frac` reads above 1.
        # The fused-history lower bound and the measured DRAM traffic say what HBM really sees.
        "fused_bound_bytes_per_launch": fused_history_bytes(timed) / max(hist_launches, 1),
        # The ceiling that does bind facet-dominated decks: every facet, census and death is one
        # red.global.add.f64 into the tally, and the B200 L2 retires 1.97e11 of those per second
        # (tools/microbench/red_rate.cu, profiles/r01/red_rate_v2.txt).
        "atomic_bound": {
            "achieved": sum(r.facets + r.census + r.deaths for r in timed) / max(kernel_ns, 1) * 1e9,
            "peak": RED_PEAK_PER_S, "unit": "fp64 reductions/s",
            "frac": sum(r.facets + r.census + r.deaths for r in timed) / max(kernel_ns, 1) * 1e9
            / RED_PEAK_PER_S,
            "peak_source": "measured, profiles/r01/red_rate_v2.txt"},
        "note": "k_history is bound by L2 FP64 atomics (facets) and FP64 issue (collisions), "
                "not by HBM (DESIGN.md 5)",
    }
    line = {
        "metric": METRIC, "value": events_all / (elapsed_ms / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{deck.name}.params: {d.nx}x{d.ny} mesh, {per_gpu} particles "
                               f"per GPU ({d.nparticles} global), {d.iterations} timesteps "
                               "per step, fresh bank each step",
                   "l2_policy": "inputs larger than L2 (density + tally = "
                                f"{2 * ncells * 8 / 2**20:.0f} MiB, random access)",
                   "parallelism": f"particle-sharded x{world}, NCCL all-reduce of the tally "
                                  "delta per timestep" if world > 1 else "single GPU",
                   "options": args.opts or "defaults (pipeline=1,fast_div=1,tile_shift=8,"
                                           "length_bins=512)",
                   "events_per_step": events_all / args.steps,
                   "tally_sum": tally_sum},
        "roofline": roofline,
        "clocks": clocks,
        "gpu_launches": launches_all,
    }
    if e2e:
        line["e2e"] = {"value": e2e_events_all / e2e_seconds, "unit": UNIT,
                       "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "ms_per_step": 1e3 * e2e_seconds / args.steps,
                       "history_kernel_ms_per_step": e2e["hist_ms"],
                       "sort_phase_ms_per_step": e2e["sort_ms"],
                       "tally_sum": e2e["tally_sum"], "live_particles_out": e2e["live_out"],
                       "pipeline": "double-buffered: upload of step i+1 and download of step "
                                   "i-1 overlap the transport of step i"}
    if world == 1 and not args.no_cpu_baseline:
        try:
            res = reference_rate(args.deck, target_seconds=15.0)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # the baseline is a reported number, never a gate
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": f"unavailable: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run "
                 "(one process per GPU)")
    run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
