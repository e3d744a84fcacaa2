import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_oracle import ORACLE_SO, build as build_oracle  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle (test infrastructure) behind the same C ABI."""
    from beamfoam_b200 import load_library
    return load_library(build_oracle())


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; the test fails (not skips) if it cannot be loaded."""
    from beamfoam_b200 import load_library
    from beamfoam_b200 import build as _build
    _build.build()
    lib = load_library()
    assert lib.bf_backend() == b"cuda-sm100a"
    return lib


@pytest.fixture(scope="session")
def oracle_lib_path():
    """Path of the oracle library, for hosts that dlopen() an implementation of the ABI themselves."""
    return build_oracle()
