import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (oracle/): the checker, never the product path."""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def cabi():
    """The product C-ABI (libpacmann_cuda.so) through ctypes; raises if the .so is missing."""
    from pacmann_b200 import cabi as c
    c.lib()
    return c
