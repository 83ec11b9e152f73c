import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer statistical checks")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native artefacts once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
