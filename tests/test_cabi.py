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
