"""The kernels' sin/cos (a transliteration of glibc's FMA-variant double sin/cos on glibc's
own table, nb_sincos.cuh) against the host libm, bit for bit, using the host build of the same
source. inject_particles (omp3/neutral.c:611-614) is the only caller: theta = 2*pi*r with r in
(0, 1]; the sweep covers every branch below glibc's huge-argument reduction."""
import ctypes as C
import math
import os

import numpy as np

_dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _sweep(lib, xs):
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    bad_x = C.c_double(0.0)
    bad = lib.nb200_selftest_host_sincos(_ptr(xs), len(xs), C.byref(bad_x))
    assert bad == 0, f"{bad} of {len(xs)} arguments differ from libm, e.g. x={bad_x.value!r}"


def test_host_sincos_matches_libm_on_the_inject_domain(lib):
    rng = np.random.default_rng(20261017)
    r = (rng.integers(0, 2 ** 64, size=4_000_000, dtype=np.uint64).astype(np.float64)
         * 2.0 ** -64 + 2.0 ** -65)           # the RNG's (0, 1] doubles, omp3/neutral.c:646-651
    _sweep(lib, 2.0 * math.pi * r)


def test_host_sincos_matches_libm_on_every_branch(lib):
    rng = np.random.default_rng(7)
    n = 1_000_000
    edges = np.array([2.0 ** -27, 2.0 ** -26, 0.126, 0.855469, 2.426265, math.pi / 2, math.pi,
                      1.5 * math.pi, 2.0 * math.pi, 105414335.0, 1.0e-300, 0.0])
    near = np.concatenate([np.nextafter(edges, np.inf), np.nextafter(edges, -np.inf), edges])
    xs = np.concatenate([
        (rng.random(n) - 0.5) * 20.0,                      # both signs, several periods
        np.ldexp(rng.random(n), -rng.integers(0, 40, n)),  # small arguments, the Taylor branch
        (rng.random(n) - 0.5) * 2.0e8,                     # up to the reduction limit
        np.arange(1, 2001) * (math.pi / 2),                # next to the multiples of pi/2
        near, -near,
    ])
    # glibc switches to its huge-argument reduction at high word 0x419921FB (~1.054e8)
    hi = (xs.view(np.uint64) >> np.uint64(32)) & np.uint64(0x7FFFFFFF)
    xs = xs[hi < 0x419921FB]
    _sweep(lib, xs)


def test_host_hooks_agree_with_ctypes_libm(lib):
    libm = C.CDLL("libm.so.6")
    for f in (libm.sin, libm.cos):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    for x in (0.5, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 2.0 * math.pi, 1e-9, 0.1259, 0.1261):
        assert lib.nb200_host_sin(x).hex() == libm.sin(x).hex()
        assert lib.nb200_host_cos(x).hex() == libm.cos(x).hex()


def test_sincos_table_matches_this_libm():
    """The committed table is the one inside the libm the oracle runs against."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "gen_sc", os.path.join(root, "tools", "gen_glibc_sincos_table.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    vals = mod.extract(mod.find_libm())
    inc = open(os.path.join(root, "neutral_b200", "csrc", "glibc_sincos_table.inc")).read()
    body = inc[inc.index("*/") + 2:]
    committed = [float.fromhex(t) for t in body.replace("\n", " ").split(",") if t.strip()]
    assert len(committed) == len(vals) == 440
    assert all(a == b for a, b in zip(committed, vals))
