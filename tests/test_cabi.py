"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/neutral_b200.h declares; compute entry points refuse to run without a device."""
import ctypes as C
import os
import re

import pytest

from neutral_b200.host import EXPORTED_SYMBOLS, NeutralB200Error

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "neutral_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)  # preprocessor lines are not prototypes
    text = re.sub(r"typedef struct \{.*?\} \w+;", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    return sorted({n for n in names if n not in ("defined", "sizeof")})


def test_every_declared_symbol_is_exported(lib):
    declared = _declared_functions()
    assert set(declared) == set(EXPORTED_SYMBOLS), set(declared) ^ set(EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version(lib):
    assert lib.nb200_abi_version() == 2


def test_no_cpu_fallback(lib):
    """Without a device the extension entry points return an error (never compute)."""
    if lib.nb200_device_count() > 0:
        pytest.skip("a GPU is present")
    assert lib.nb200_synchronize() != 0
    assert b"no CPU fallback" in lib.nb200_last_error()
    from neutral_b200.decks import build_problem
    from neutral_b200.host import Simulation
    with pytest.raises(NeutralB200Error):
        Simulation(build_problem("stream_small"))


def test_product_never_touches_the_oracle():
    """Nothing under neutral_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "neutral_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                text = open(os.path.join(dirpath, f)).read()
                hit = re.search(r"import\s+oracle|from\s+oracle|oracle[/.\\]|neutral_oracle|"
                                r"libneutral_omp3|_ref", text)
                assert hit is None, (os.path.join(dirpath, f), hit.group(0))


def test_options_are_validated_without_a_device(lib):
    """Option names and ranges are checked on the host (no GPU involved): unknown names and
    out-of-range values give NB200_BAD_OPTION, legitimate negative values pass."""
    from neutral_b200.host import NB200_BAD_OPTION
    assert lib.nb200_set_option(b"no_such_option", 1) == NB200_BAD_OPTION
    assert b"unknown option" in lib.nb200_last_error()
    assert lib.nb200_set_option(b"tile_shift", 99) == NB200_BAD_OPTION
    assert lib.nb200_set_option(b"ngpus", 1000) == NB200_BAD_OPTION
    assert lib.nb200_get_option(b"no_such_option") == NB200_BAD_OPTION
    prev = lib.nb200_set_option(b"tile_shift", -1)
    assert prev != NB200_BAD_OPTION and lib.nb200_get_option(b"tile_shift") == -1
    assert lib.nb200_set_option(b"tile_shift", prev) == -1
    for name in (b"tally_reduce_every", b"collective", b"reduce_ctas", b"host_mirror",
                 b"headroom_pct", b"defer_finish", b"length_bins"):
        assert lib.nb200_get_option(name) != NB200_BAD_OPTION, name


def test_group_entry_points_refuse_without_a_device(lib):
    import ctypes as C
    if lib.nb200_device_count() > 0:
        pytest.skip("a GPU is present")
    blob = (C.c_char * 256)()
    assert lib.nb200_mp_init(2, 0, 1024, blob) == -1
    assert lib.nb200_tally_sync(None) == -1
    assert lib.nb200_mp_finalize() == 0  # nothing to tear down
    rate = C.c_double()
    assert lib.nb200_microbench_red(3, 16 << 20, 10, C.byref(rate)) == -1
    assert b"no CPU fallback" in lib.nb200_last_error()


def test_staggered_dispatch_is_a_permutation_of_the_live_groups(lib):
    """The event loop's CTA -> bank-group map under the stagger_* options (history.cu:
    dispatch_group, evaluated on the host by the library itself): every group of the live prefix
    is visited exactly once whatever the arguments, the map is the identity when the option is
    off, when the colliders are too few or exceed the first wave, and beyond the live prefix;
    where it applies, the colliders that start with the launch come first and the delayed ones
    sit behind the requested share of the streamer groups."""
    import random
    f = lib.nb200_selftest_dispatch_group
    f.argtypes = [C.c_uint] * 3 + [C.c_int] * 3
    f.restype = C.c_uint
    rng = random.Random(7)
    cases = [(1_000_000, 55_169, 40, 50, 35), (916_929, 56_606, 55, 35, 40), (1000, 0, 40, 50, 0),
             (129, 128, 40, 50, 0), (128 * 900, 128 * 889, 40, 50, 0), (5000, 5000, 40, 50, 0)]
    for _ in range(200):
        n_live = rng.randrange(0, 300_000)
        cases.append((n_live, rng.randrange(0, n_live + 1), rng.randrange(0, 96),
                      rng.randrange(0, 101), rng.randrange(0, 200)))
    for n_live, n_coll, at, share, mn in cases:
        groups = (n_live + 127) // 128 + 3  # the grid covers the host's upper bound, not n_live
        image = [f(b, n_live, n_coll, at, share, mn) for b in range(groups)]
        assert sorted(image) == list(range(groups)), (n_live, n_coll, at, share, mn)
        full, nc = n_live // 128, n_coll // 128
        applies = at > 0 and nc < full and nc <= 148 * 6 and n_coll * 1000 >= mn * n_live
        if not applies:
            assert image == list(range(groups))
            continue
        assert image[full:] == list(range(full, groups))  # ragged tail and dead groups keep their place
        delayed = nc * share // 100
        first = nc - delayed
        at_group = first + (full - nc) * at // 100
        assert image[:first] == list(range(first))
        assert image[at_group:at_group + delayed] == list(range(first, nc))
    # the csp timestep of profiles/r02/experiments/warp_trace_stagger40_csp_step6.txt
    assert f(0, 916_929, 56_606, 0, 50, 40) == 0 and f(300, 916_929, 56_606, 40, 50, 40) == 300 + 221
