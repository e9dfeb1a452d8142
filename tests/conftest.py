import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


SMALL = dict(logN=10, L=6, dnum=3)


@pytest.fixture(scope="session")
def small_oracle():
    from oracle.oracle import Oracle
    return Oracle(**SMALL)


@pytest.fixture(scope="session")
def ref_oracle():
    """Reference parameters (FHEController.cpp:6-35): N=2^15, 28 Q limbs, dnum 4."""
    from oracle.oracle import Oracle
    return Oracle(logN=15, L=28, dnum=4)
