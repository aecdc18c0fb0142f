import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the product library and the CPU checkers once per session."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """GPU and CPU tests alike need the in-tree shared libraries; build them once if a fresh checkout lacks them."""
    lib = os.path.join(ROOT, "lbm_b200", "liblbm_b200.so")
    orc = os.path.join(ROOT, "oracle", "liblbm_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()
