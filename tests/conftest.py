import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SMALL_DECKS = ["csp_small", "split_small", "scatter_small", "stream_small", "mixed_small"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """The plain-C oracle port (test infrastructure)."""
    from oracle.oracle import OraclePort
    return OraclePort()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference omp3 kernel set, when its prebuilt library is around."""
    from oracle.oracle import ReferenceOmp3
    if not ReferenceOmp3.available():
        pytest.skip("oracle/_ref/libneutral_omp3.so not built and /root/reference absent")
    return ReferenceOmp3()


@pytest.fixture(scope="session")
def lib():
    """libneutral_b200.so through ctypes (built in-tree; loading needs no GPU)."""
    from neutral_b200.host import load_library
    return load_library(build=True)


@pytest.fixture(scope="session")
def gpu_lib(lib):
    if lib.nb200_device_count() <= 0:
        pytest.fail("a gpu-marked test ran without a CUDA device")
    return lib
