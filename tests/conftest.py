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
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def native():
    """The CUDA library on a box with a usable sm_100 device; the GPU tests fail (not skip)
    when it is missing, so a silent fallback can never pass them."""
    from erlvectordb_b200 import _native as N
    L = N.lib()
    rc = L.evdb_init(None, 0)
    assert rc == 0, f"evdb_init: {L.evdb_strerror(rc).decode()}"
    return N


def fromhex(x):
    return float.fromhex(x)
