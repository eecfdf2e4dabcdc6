import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robot-parallel-motion-planning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracles():
    """Builds (if needed) and loads the CPU oracles: the checker, never the product."""
    from oracle import build_oracle
    from oracle.gmt_oracle import OraclePort, RefHost
    build_oracle.build_port()
    try:
        build_oracle.build_ref()
    except Exception:
        pass
    out = {"libm": OraclePort("libm"), "cuda": OraclePort("cuda")}
    out["ref"] = RefHost() if RefHost.available() else None
    return out


@pytest.fixture(scope="session")
def gmt_lib():
    """The product library (built in-tree by __graft_entry__.build / build.py)."""
    import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import build as gmt_build
        gmt_build.build()
    return _lib.load()
