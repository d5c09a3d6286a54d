import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the native artefacts exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    return g


@pytest.fixture(scope="session")
def hostemu(built):
    from tests import helpers
    return helpers.HostEmu()
