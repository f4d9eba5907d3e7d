import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Build the product library and the checkers if they are not there yet (CPU-only is fine:
    nvcc cross-compiles; the GPU box receives the prebuilt files)."""
    need = [os.path.join(ROOT, "canvas_ity_b200", "libcanvas_b200.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


def has_gpu():
    from canvas_ity_b200 import _native
    try:
        return _native.load().cb200_device_count() > 0
    except Exception:
        return False
