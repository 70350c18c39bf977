import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (test infrastructure)."""
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def pymodel():
    from oracle import pymodel as M

    return M


@pytest.fixture(scope="session")
def hodor():
    """The product, bound to cuda:0.  Fails (not skips) when the library or the GPU is missing."""
    import hodor_b200 as H

    H.init(0)
    return H


FIELDS = [0, 1, 2]
FIELD_IDS = ["bls12_381_fr", "bn254_fr", "stark252"]
