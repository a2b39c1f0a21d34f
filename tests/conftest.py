import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
