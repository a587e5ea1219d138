import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port_oracle():
    from oracle.oracle import PortOracle

    return PortOracle()


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.oracle import RefOracle

    if not RefOracle.available():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference)")
    return RefOracle()


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import best_oracle

    return best_oracle()


@pytest.fixture(scope="session")
def ctx():
    import combblas_b200 as cb

    c = cb.Context(0)
    yield c
    c.close()
