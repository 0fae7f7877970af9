import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """Builds (if needed) and loads the sm_100a library; CPU-only boxes can still load it."""
    from saro_gs_b200 import build as native_build
    native_build.build(verbose=False)
    from saro_gs_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def has_golden(name):
    return os.path.exists(golden_path(name))
