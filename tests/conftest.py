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


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not errored) where there is no CUDA device; the product itself has no CPU path."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
    import shutil
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")) and os.path.exists(os.path.join(PKG, "topkrec", "libtopkrec.so")):
        return          # no compiler on this box: use the library that travelled with the tree
    g.build()
