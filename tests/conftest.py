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


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: skipped (not failed) on a box without one.  The
    refusal to run without a GPU is itself tested by the unmarked test_no_cpu_fallback."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import __graft_entry__ as g
        g.build()
        from exchanger_b200 import capi
        have = capi.load_library().exb200_device_count() > 0
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device visible (exb200_device_count() == 0)")
        for it in gpu_items:
            it.add_marker(skip)
