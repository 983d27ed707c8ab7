"""The engine around the kernels, on the GPU: pipelined timesteps, banks sharded over several
GPUs inside the library (peer-memory reduce-scatter kernel and the NCCL flavour), the exported
`validate`, the drop-in binary built from the reference's unmodified main.c, option
validation, in-place table edits, the host mirror for visit_dump decks and the head-room for
produced particles. Everything goes through the C ABI; the oracle is only the checker."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from neutral_b200.bank import ALL_FIELDS, F64_FIELDS, I32_FIELDS, HostBank, ParticleSoA
from neutral_b200.decks import build_problem
from neutral_b200.host import NB200_BAD_OPTION, DeviceArray, Simulation, _check
from test_gpu_parity import tally_close

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN_DIR = os.path.join(ROOT, "build", "run", "neutral")
DROPIN = os.path.join(RUN_DIR, "neutral.b200")


def _golden_full():
    with open(os.path.join(ROOT, "tests", "golden", "full_decks.json")) as f:
        return json.load(f)


def _capture_stdout(fn):
    """Runs fn() with file descriptor 1 redirected to a pipe file (C printf included)."""
    import tempfile
    C.CDLL(None).fflush(None)
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            C.CDLL(None).fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


@pytest.mark.parametrize("deck", ["csp_small", "mixed_small"])
def test_pipelined_timesteps_equal_one_at_a_time(gpu_lib, deck):
    """Up to three timesteps enqueued before the oldest is collected (defer_finish): the same
    counts per timestep, the same bank bit for bit, the same tally."""
    prob = build_problem(deck)
    a = Simulation(prob)
    a.inject()
    ra = a.run()
    b = Simulation(prob)
    b.inject()
    rb = b.run_pipelined(depth=3)
    key = lambda r: (r.facets, r.collisions, r.processed, r.census, r.deaths)
    assert [key(r) for r in ra] == [key(r) for r in rb]
    assert sum(a.bank_to_host().bit_equal(b.bank_to_host()).values()) == 0
    assert np.array_equal(a.counters_to_host(), b.counters_to_host())
    assert tally_close(a.tally_to_host(), b.tally_to_host())
    a.free()
    b.free()


def test_bank_operations_are_refused_while_a_timestep_is_pending(gpu_lib):
    prob = build_problem("stream_small")
    sim = Simulation(prob, per_particle_counters=False)
    sim.inject()
    other = Simulation(prob, per_particle_counters=False)
    other.inject()
    sim.step(1, defer=True)
    assert gpu_lib.nb200_bank_pending(sim.bank) == 1
    host = HostBank.empty(sim.count)
    st = host.as_struct()
    assert gpu_lib.nb200_bank_download(sim.bank, C.byref(st)) == -4
    assert gpu_lib.nb200_bank_copy(sim.bank, other.bank) == -4
    assert gpu_lib.nb200_bank_export(sim.bank) == -4
    assert gpu_lib.nb200_set_stream(None) == -4
    assert b"nb200_solve_finish" in gpu_lib.nb200_last_error()
    r = sim.step_finish()
    assert r.processed == sim.count
    assert gpu_lib.nb200_bank_download(sim.bank, C.byref(st)) == 0
    assert gpu_lib.nb200_set_stream(None) == 0
    sim.free()
    other.free()


def test_options_are_validated(gpu_lib):
    assert gpu_lib.nb200_set_option(b"no_such_option", 1) == NB200_BAD_OPTION
    assert gpu_lib.nb200_set_option(b"tile_shift", 99) == NB200_BAD_OPTION
    assert b"tile_shift" in gpu_lib.nb200_last_error()
    assert gpu_lib.nb200_set_option(b"length_bins", -5) == NB200_BAD_OPTION
    assert gpu_lib.nb200_get_option(b"tile_shift") == 8
    prev = gpu_lib.nb200_set_option(b"tile_shift", -1)  # a legitimate negative value
    assert prev == 8 and gpu_lib.nb200_set_option(b"tile_shift", prev) == -1


def test_a_sort_key_that_would_overflow_gets_coarser_tiles(gpu_lib):
    """tile_shift = 0 with 512 length bins on the 4000 x 4000 mesh asks for 2.4e10 histogram
    bins (the int overflow of ADVICE r1): the sort falls back to coarser tiles and the run is
    the default configuration's, count for count."""
    prob = build_problem("csp", nparticles=20_000)
    ref_sim = Simulation(prob, per_particle_counters=False)
    ref_sim.inject()
    want = [(r.facets, r.collisions) for r in ref_sim.run(3)]
    bank_want = ref_sim.bank_to_host()
    ref_sim.free()
    sim = Simulation(prob, per_particle_counters=False)
    sim.inject()
    assert gpu_lib.nb200_bank_set_option(sim.bank, b"tile_shift", 0) == 8
    got = [(r.facets, r.collisions) for r in sim.run(3)]
    assert got == want
    assert sum(sim.bank_to_host().bit_equal(bank_want).values()) == 0
    sim.free()
    assert gpu_lib.nb200_get_option(b"tile_shift") == 8  # the override was the bank's own


def test_per_bank_options_do_not_leak(gpu_lib, port):
    """Two banks with different kernel configurations step side by side."""
    prob = build_problem("mixed_small")
    d = prob.deck
    a = Simulation(prob, per_particle_counters=False)
    a.inject()
    b = Simulation(prob, per_particle_counters=False)
    b.inject()
    for k, v in dict(pipeline=0, fast_div=0).items():
        assert gpu_lib.nb200_bank_set_option(b.bank, k.encode(), v) != NB200_BAD_OPTION
    bank = port.inject(prob)
    tally = np.zeros(d.nx * d.ny)
    for tt in range(1, d.iterations + 1):
        want = port.step(prob, bank, tt, tally)
        ra, rb = a.step(tt), b.step(tt)
        assert (ra.facets, ra.collisions, ra.processed) == want
        assert (rb.facets, rb.collisions, rb.processed) == want
        assert ra.launches > rb.launches  # the direct kernel has no sort phase
    assert sum(a.bank_to_host().bit_equal(bank).values()) == 0
    assert sum(b.bank_to_host().bit_equal(bank).values()) == 0
    a.free()
    b.free()


def test_tables_edited_in_place_between_timesteps(gpu_lib, port):
    """The caller may change a table where it lies between two timesteps (the reference would
    simply read the new values): same pointers, same length, another energy grid. The
    same-grid shortcut is decided on the device every step, so nothing stale survives
    (capi.cu:340)."""
    prob = build_problem("mixed_small")
    d = prob.deck
    sim = Simulation(prob)
    sim.inject()
    bank = port.inject(prob)
    tally = np.zeros(d.nx * d.ny)
    want = port.step(prob, bank, 1, tally)
    got = sim.step(1)
    assert (got.facets, got.collisions, got.processed) == want
    # capture table on a grid of its own: every key moved by a few ulps-worth, values scaled
    keys, vals = prob.cs_absorb
    new_keys = np.ascontiguousarray(keys * (1.0 + 1e-7))
    new_vals = np.ascontiguousarray(vals * 0.5)
    assert np.all(np.diff(new_keys) > 0)
    prob.cs_absorb = (new_keys, new_vals)
    sim._cs_arrays[2].upload(new_keys)   # in place: same device pointers
    sim._cs_arrays[3].upload(new_vals)
    for tt in (2, 3):
        want = port.step(prob, bank, tt, tally)
        got = sim.step(tt)
        assert (got.facets, got.collisions, got.processed) == want, tt
        assert sum(sim.bank_to_host().bit_equal(bank).values()) == 0, tt
    assert tally_close(sim.tally_to_host(), tally)
    sim.free()


def test_exported_validate_prints_the_reference_verdict(gpu_lib, tmp_path):
    """The exported `validate` (neutral_interface.h:35-36) on the full csp deck: the reference's
    printed lines, PASSED against problems/neutral.tests, and the machine-readable results."""
    prob = build_problem("csp")
    d = prob.deck
    sim = Simulation(prob, per_particle_counters=False)
    sim.inject()
    sim.run_pipelined()
    out_json = tmp_path / "results.json"
    os.environ["NB200_RESULTS_JSON"] = str(out_json)
    cwd = os.getcwd()
    os.chdir(ROOT)  # validate opens problems/neutral.tests relative to the run directory
    try:
        text = _capture_stdout(lambda: gpu_lib.validate(d.nx, d.ny, b"problems/csp.params", 0,
                                                        sim.tally.ptr))
    finally:
        os.chdir(cwd)
        del os.environ["NB200_RESULTS_JSON"]
    assert "Final global_energy_tally" in text
    assert "Expected 1.121870290714e+07" in text
    assert "PASSED validation." in text
    res = json.loads(out_json.read_text())
    assert res["verdict"] == "PASSED" and res["relative_error"] < 1e-3
    g = _golden_full()["csp"]
    assert abs(res["tally_total"] - g["tally_sum"]) <= 1e-9 * g["tally_sum"]
    steps = res["steps"][-d.iterations:]
    assert [[s["facets"], s["collisions"]] for s in steps] == g["counts"]
    # a deck without an entry: the reference's warning
    text = _capture_stdout(lambda: gpu_lib.validate(d.nx, d.ny, b"problems/split.params", 0,
                                                    sim.tally.ptr))
    assert "could NOT validate" in text
    sim.free()


def _run_dropin(deck, env_extra=None, timeout=300):
    if not os.path.exists(DROPIN):
        pytest.skip("build/run/neutral/neutral.b200 has not been built (it needs the reference's "
                    "main.c: __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, **(env_extra or {}))
    env.pop("OMP_NUM_THREADS", None)
    res = subprocess.run([DROPIN, f"problems/{deck}.params"], cwd=RUN_DIR, env=env,
                         capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    facets = [int(x) for x in re.findall(r"^Facets\s+(\d+)", res.stdout, flags=re.M)]
    colls = [int(x) for x in re.findall(r"^Collisions\s+(\d+)", res.stdout, flags=re.M)]
    return res.stdout, [[f, c] for f, c in zip(facets, colls)]


def test_dropin_binary_runs_the_reference_driver(gpu_lib, tmp_path):
    """`make KERNELS=b200`: the reference's unmodified main.c + neutral_data.c linked against
    libneutral_b200.so, run like the reference (cwd = the neutral directory). Per-timestep
    Facets / Collisions lines equal the reference omp3 run's, validate PASSES."""
    out_json = tmp_path / "dropin.json"
    text, counts = _run_dropin("csp", {"NB200_RESULTS_JSON": str(out_json)})
    g = _golden_full()["csp"]
    assert counts == g["counts"]
    assert "PASSED validation." in text
    res = json.loads(out_json.read_text())
    assert res["verdict"] == "PASSED" and res["deck"] == "problems/csp.params"
    assert [s["gpus"] for s in res["steps"]] == [1] * len(g["counts"])


def _need_gpus(lib, n):
    if lib.nb200_device_count() < n:
        pytest.skip(f"needs {n} GPUs, {lib.nb200_device_count()} visible")


@pytest.mark.parametrize("reduce_every", [0, 1, 2])
@pytest.mark.parametrize("collective", [1, 0])
@pytest.mark.parametrize("deck", ["mixed_small", "csp_small"])
def test_bank_sharded_over_gpus_inside_the_library(gpu_lib, port, deck, collective, reduce_every):
    """One process, several GPUs (option ngpus): same calls as a single-GPU run; the shards
    replay their particles bit for bit, the counts are the sums, the per-particle counters land
    in the caller's arrays, and the tally (reduce-scattered into owned slices every timestep,
    gathered when somebody looks) is the single-GPU tally within summation order."""
    _need_gpus(gpu_lib, 2)
    ngpus = min(gpu_lib.nb200_device_count(), 4)
    prob = build_problem(deck)
    d = prob.deck
    prev = gpu_lib.nb200_set_option(b"collective", collective)
    prev_every = gpu_lib.nb200_set_option(b"tally_reduce_every", reduce_every)
    try:
        sim = Simulation(prob, ngpus=ngpus, per_particle_counters=bool(collective))
        sim.inject()
        assert gpu_lib.nb200_bank_gpus(sim.bank) == ngpus
        bank = port.inject(prob)
        assert sum(sim.bank_to_host().bit_equal(bank).values()) == 0, "inject differs"
        tally = np.zeros(d.nx * d.ny)
        ctr = np.zeros((3, len(bank)), dtype=np.uint64)
        for tt in range(1, d.iterations + 1):
            want = port.step(prob, bank, tt, tally, counters=ctr)
            got = sim.step(tt)
            assert (got.facets, got.collisions, got.processed) == want, tt
            assert sum(sim.bank_to_host().bit_equal(bank).values()) == 0, tt
            if tt in (1, d.iterations):  # looking at the tally gathers the owned slices
                assert tally_close(sim.tally_to_host(), tally), tt
        if collective:
            assert np.array_equal(sim.counters_to_host(), ctr)
        assert tally_close(sim.tally_to_host(), tally)
        sim.free()
    finally:
        gpu_lib.nb200_set_option(b"collective", prev)
        gpu_lib.nb200_set_option(b"tally_reduce_every", prev_every)


def test_sharded_full_deck_matches_the_reference(gpu_lib):
    """split at full size on every visible GPU of one process, pipelined: per-timestep counts of
    the reference run, final bank bit-identical, tally total and block sums within 1e-9."""
    import hashlib
    _need_gpus(gpu_lib, 2)
    ngpus = gpu_lib.nb200_device_count()
    g = _golden_full()["split"]
    prob = build_problem("split")
    d = prob.deck
    sim = Simulation(prob, ngpus=ngpus, per_particle_counters=False)
    sim.inject()
    res = sim.run_pipelined()
    assert [[r.facets, r.collisions] for r in res] == g["counts"]
    bank = sim.bank_to_host()
    got = {k: hashlib.sha256(np.ascontiguousarray(bank.arrays[k]).tobytes()).hexdigest()
           for k in ALL_FIELDS}
    assert got == g["final_hashes"]
    tally = sim.tally_to_host()
    assert abs(float(tally.sum()) - g["tally_sum"]) <= 1e-9 * g["tally_sum"]
    sim.free()


def test_dropin_binary_on_several_gpus(gpu_lib):
    """NB200_NGPUS=<n> ./neutral.b200 problems/split.params: the unmodified driver, n GPUs."""
    _need_gpus(gpu_lib, 2)
    n = gpu_lib.nb200_device_count()
    text, counts = _run_dropin("split", {"NB200_NGPUS": str(n)})
    assert counts == _golden_full()["split"]["counts"]
    assert f"sharded over {n} GPUs" in text
    tally = float(re.search(r"Final global_energy_tally (\S+)", text).group(1))
    want = _golden_full()["split"]["tally_sum"]
    assert abs(tally - want) <= 1e-9 * want


def test_host_mirror_for_visit_dump_decks(gpu_lib, port):
    """Option host_mirror: the handle's 11 pointers are HOST arrays in injection order, kept up
    to date, and the handle is an array of one such struct per particle - what the reference's
    plot_particle_density (main.c:169-187, `&local_particles[ii]` under -DSoA) indexes."""
    prob = build_problem("csp_small", nparticles=5_000)
    prev = gpu_lib.nb200_set_option(b"host_mirror", 1)
    try:
        sim = Simulation(prob, per_particle_counters=False)
        sim.inject()
    finally:
        gpu_lib.nb200_set_option(b"host_mirror", prev)
    bank = port.inject(prob)
    tally = np.zeros(prob.deck.nx * prob.deck.ny)

    def mirror(index):
        views = C.cast(sim.bank, C.POINTER(ParticleSoA))
        v = views[index]
        out = {}
        for k in F64_FIELDS:
            out[k] = np.ctypeslib.as_array(getattr(v, k), shape=(sim.count,)).copy()
        for k in I32_FIELDS:
            out[k] = np.ctypeslib.as_array(getattr(v, k), shape=(sim.count,)).copy()
        return HostBank(out)

    assert sum(mirror(0).bit_equal(bank).values()) == 0
    port.step(prob, bank, 1, tally)
    sim.step(1)
    assert sum(mirror(0).bit_equal(bank).values()) == 0
    assert sum(mirror(sim.count - 1).bit_equal(bank).values()) == 0  # every copy of the struct
    sim.free()


def test_headroom_takes_produced_particles(gpu_lib, port):
    """The reference allocates twice the bank (omp3/neutral.c:570) for particles a physics
    extension would produce; option headroom_pct does the same and nb200_bank_append fills it:
    appended particles get the next global indices and transport like injected ones."""
    prob = build_problem("mixed_small")
    d = prob.deck
    n_first = d.nparticles - 1000
    prev = gpu_lib.nb200_set_option(b"headroom_pct", 100)
    try:
        first = port.inject(prob, 0, n_first)
        sim = Simulation(build_problem("mixed_small"), per_particle_counters=False)
        sim.count = n_first
        sim.load_bank(first)
    finally:
        gpu_lib.nb200_set_option(b"headroom_pct", prev)
    assert gpu_lib.nb200_bank_capacity(sim.bank) == 2 * n_first
    tally = np.zeros(d.nx * d.ny)
    pf, pc, pp = port.step(prob, first, 1, tally)
    r = sim.step(1)
    assert (r.facets, r.collisions, r.processed) == (pf, pc, pp)
    extra = port.inject(prob, n_first, 1000)  # particles n_first .. n_first + 999
    st = extra.as_struct()
    _check(gpu_lib.nb200_bank_append(sim.bank, C.byref(st), 1000), "bank_append")
    assert gpu_lib.nb200_bank_size(sim.bank) == d.nparticles
    sim.count = d.nparticles
    grown = HostBank({k: np.concatenate([first.arrays[k], extra.arrays[k]]) for k in ALL_FIELDS})
    for tt in (2, 3):
        want = port.step(prob, grown, tt, tally)
        got = sim.step(tt)
        assert (got.facets, got.collisions, got.processed) == want, tt
        assert sum(sim.bank_to_host().bit_equal(grown).values()) == 0, tt
    assert tally_close(sim.tally_to_host(), tally)
    too_many = port.inject(prob, 0, 2 * n_first)
    st = too_many.as_struct()
    assert gpu_lib.nb200_bank_append(sim.bank, C.byref(st), 2 * n_first) == -5
    sim.free()


def test_red_microbenchmark_reports_a_rate(gpu_lib):
    rate = C.c_double(0.0)
    assert gpu_lib.nb200_microbench_red(3, 16 << 20, 50, C.byref(rate)) == 0
    assert 1e10 < rate.value < 1e12


def test_dropin_visit_dump_deck(gpu_lib, port, tmp_path):
    """A deck with `visit_dump 1` (main.c:91-94,129-139,169-200): the driver reads the bank on
    the host and dumps the tally every timestep. With NB200_HOST_MIRROR=1 the b200 kernel set
    serves both: the particle-density plot of the injected bank counts every particle, and
    the dumped tallies equal the reference omp3 binary's dumps cell by cell (1e-10)."""
    if not os.path.exists(DROPIN):
        pytest.skip("build/run/neutral/neutral.b200 has not been built")
    ref_dir = os.path.join(ROOT, "oracle", "_ref", "run", "neutral")
    ref_exe = os.path.join(ref_dir, "neutral.omp3")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/run/neutral/neutral.omp3 has not been built")
    deck = "problems/small/visit_small.params"

    def run(exe, src_dir, where, env_extra):
        os.makedirs(where)
        for name in ("elastic_scatter.cs", "capture.cs"):
            os.symlink(os.path.join(src_dir, name), os.path.join(where, name))
        os.symlink(os.path.join(ROOT, "problems"), os.path.join(where, "problems"))
        with open(os.path.join(os.path.dirname(where), "arch.params"), "w") as f:
            f.write(open(os.path.join(ROOT, "archlite", "arch.params")).read())
        res = subprocess.run([exe, deck], cwd=where, env=dict(os.environ, **env_extra),
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
        return res.stdout

    gpu_dir = str(tmp_path / "gpu" / "neutral")
    cpu_dir = str(tmp_path / "cpu" / "neutral")
    out_gpu = run(DROPIN, RUN_DIR, gpu_dir, {"NB200_HOST_MIRROR": "1", "MALLOC_PERTURB_": "255"})
    out_cpu = run(ref_exe, ref_dir, cpu_dir, {"OMP_NUM_THREADS": "4", "MALLOC_PERTURB_": "255"})
    grab = lambda text, key: re.findall(rf"^{key}\s+(\d+)", text, flags=re.M)
    assert grab(out_gpu, "Facets") == grab(out_cpu, "Facets")
    assert grab(out_gpu, "Collisions") == grab(out_cpu, "Collisions")
    # the particle-density plot of the injected bank: one count per particle in its cell. The
    # reference accumulates into an UNINITIALISED malloc (main.c:171-172), so cells carry
    # whatever the allocator left there (seen on the GPU box: three denormals, 2.6e-319 ...):
    # both binaries run with glibc's MALLOC_PERTURB_=255, which fills malloc'ed memory with
    # ~0xff = 0x00, and the counts are compared with the oracle's injected bank.
    parts = np.fromfile(os.path.join(gpu_dir, "particles1.dat"))
    bank = port.inject(build_problem("visit_small"))
    want = np.bincount(bank.celly.astype(np.int64) * 256 + bank.cellx, minlength=256 * 256)
    assert parts.size == 256 * 256 and parts.sum() == 6000.0
    assert np.array_equal(parts, want.astype(np.float64))
    for tt in (1, 2, 3, 4):  # ... and after every timestep: the mirror follows the bank
        a = np.fromfile(os.path.join(gpu_dir, f"particles{tt}.dat"))
        b = np.fromfile(os.path.join(cpu_dir, f"particles{tt}.dat"))
        assert a.sum() == 6000.0 and np.array_equal(a, b), tt
    for tt in (1, 2, 3):
        a = np.fromfile(os.path.join(gpu_dir, f"energy{tt}.dat"))
        b = np.fromfile(os.path.join(cpu_dir, f"energy{tt}.dat"))
        assert tally_close(a, b), tt
        assert "DATA_SIZE: 256 256 1" in open(os.path.join(gpu_dir, f"energy{tt}.bov")).read()
