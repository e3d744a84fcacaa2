import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libbeamoracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def build_oracle():
    src = os.path.join(ROOT, "oracle", "beam_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return ORACLE_SO


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
