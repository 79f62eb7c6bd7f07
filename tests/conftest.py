import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The product libraries are built in-tree by __graft_entry__.build(); build them here if a
    test session starts on a fresh checkout (CPU box: nvcc cross-compiles)."""
    import fpmb200
    import fpmhost
    if not (os.path.exists(fpmb200.lib_path()) and os.path.exists(fpmhost.lib_path())):
        import __graft_entry__
        __graft_entry__.build()
    yield
