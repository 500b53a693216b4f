import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the product libraries and the oracle; return their paths."""
    from libmodjpeg_b200 import build as B
    from oracle import oracle_py as O

    out = B.build()
    O.build()
    return out


@pytest.fixture(scope="session")
def port(built):
    from oracle import oracle_py as O

    return O.OraclePort()


@pytest.fixture(scope="session")
def ref(built):
    from oracle import oracle_py as O

    if not O.have_reference():
        pytest.skip("oracle/_ref not built (no /root/reference here and no prebuilt copy)")
    return O.Reference()


@pytest.fixture(scope="session")
def engine(built):
    from libmodjpeg_b200 import Engine

    return Engine(int(os.environ.get("MJX_DEVICE", "0")))
