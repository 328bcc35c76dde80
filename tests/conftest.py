import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library and the oracle once per session if they are missing or stale."""
    import __graft_entry__ as ge
    ge.build()
    yield


@pytest.fixture(scope="session")
def ctx():
    import ezpz_b200 as ez
    return ez.Context(0)
