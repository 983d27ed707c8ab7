"""The kernels' log (a transliteration of glibc's FMA-variant double log on glibc's own
table) against the host libm, bit for bit, using the host build of the same source."""
import ctypes as C
import math

import numpy as np


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def test_host_log_matches_libm_bitwise(lib):
    rng = np.random.default_rng(20261017)
    n = 200_000
    xs = np.concatenate([
        rng.random(n),                         # the RNG's range
        0.93 + 0.14 * rng.random(n),           # both sides of the near-1 branch
        np.exp(-45.0 * rng.random(n)),         # down to 2^-65
        np.array([1.0, 0.9375, 1.0647, 2.0 ** -65, 2.0 ** -64 + 2.0 ** -65, 0.5, 0.75,
                  np.nextafter(1.0, 0.0), np.nextafter(1.0, 2.0), np.nextafter(0.9375, 0.0)]),
    ])
    libm = C.CDLL("libm.so.6")
    libm.log.restype = C.c_double
    libm.log.argtypes = [C.c_double]
    want = np.array([libm.log(float(x)) for x in xs])
    got = np.array([lib.nb200_host_log(float(x)) for x in xs])
    assert np.array_equal(_bits(want), _bits(got))
    assert lib.nb200_host_log(1.0) == 0.0 and math.copysign(1.0, lib.nb200_host_log(1.0)) == 1.0


def test_log_table_matches_this_libm():
    """The committed table is the one inside the libm the oracle runs against."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "gen_table", os.path.join(root, "tools", "gen_glibc_log_table.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    vals = mod.extract(mod.find_libm())
    inc = open(os.path.join(root, "neutral_b200", "csrc", "glibc_log_table.inc")).read()
    body = inc[inc.index("*/") + 2:]
    committed = [float.fromhex(t) for t in body.replace("\n", " ").split(",") if t.strip()]
    assert len(committed) == len(vals) == 274
    assert all(a == b for a, b in zip(committed, vals))
