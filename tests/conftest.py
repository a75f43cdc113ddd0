import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def core():
    """This package's host mirror of raypier.core."""
    import raypier_optics_b200.core as c
    return c


@pytest.fixture(scope="session")
def refcore():
    """The genuine reference core built into oracle/_ref (absent on the GPU box)."""
    from oracle import oracle as O
    c = O.import_reference("parity")
    if c is None:
        pytest.skip("reference not built here (oracle/_ref missing)")
    return c


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine.  No fallback: a missing librpx.so or GPU is a hard failure."""
    from raypier_optics_b200.engine import get_engine
    return get_engine(0)
