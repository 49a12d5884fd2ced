import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the C-ABI library and the C oracle are compiled."""
    import __graft_entry__ as g

    from ribotricer_b200 import _lib
    from oracle import c_oracle
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(c_oracle.LIB_PATH):
        g.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    from ribotricer_b200.engine import Engine

    eng = Engine(0)
    yield eng
    eng.close()
