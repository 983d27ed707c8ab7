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
