import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both libraries must exist; build the CPU oracle on demand (seconds)."""
    from oracle import oracle as O
    O.build()
    yield
