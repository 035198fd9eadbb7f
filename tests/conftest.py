import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C-ABI; fails loudly if the CUDA library is missing."""
    import privacy_preserving_sfm_b200 as pp
    c = pp.Context(0)
    yield c
    c.close()
