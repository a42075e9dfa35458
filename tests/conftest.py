import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    config.addinivalue_line("markers", "slow: longer statistical tests")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (nvcc cross-compiles on CPU-only boxes)."""
    from baghera_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def cfg1():
    from baghera_b200 import synth
    return synth.make_arrays(20_000, 200)
