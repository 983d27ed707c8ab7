"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl
reference`) prints ONE JSON line with the keys the driver reads, times the CPU checker with
the threads it reports - also when the launcher exported OMP_NUM_THREADS=1, as torchrun
does - and the b200 arm refuses to run N > 1 outside torchrun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT,
                          capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    res = _run(["--impl", "reference", "--deck", "csp_small", "--steps", "1", "--warmup", "0"],
               {"OMP_NUM_THREADS": "1", "NB200_REF_THREADS": "2"})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_events_per_sec"
    assert d["unit"] == "events/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 2 and cb["value"] == d["value"]
    assert "OMP threads=2" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    res = _run(["--impl", "reference", "--deck", "csp_small", "--steps", "1", "--warmup", "0"],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_multi_gpu_needs_torchrun():
    res = _run(["--gpus", "2", "--steps", "1", "--warmup", "0"])
    assert res.returncode != 0 and "torch.distributed.run" in (res.stderr + res.stdout)
