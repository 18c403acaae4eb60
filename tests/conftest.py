import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "top-k-rec_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


@pytest.fixture(scope="session")
def mini():
    return os.path.join(GOLDEN, "mini")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build the CUDA library and the C oracle once per session (nvcc cross-compiles
    without a GPU); the product has no fallback, so a failed build fails the suite."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
