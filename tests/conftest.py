import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_with_optics(oracle):
    from simc_gfortran_b200 import load_optics_fixture
    for arm in (1, 2, 3, 4, 5):
        oracle.set_optics(load_optics_fixture(arm))
    return oracle


@pytest.fixture(scope="session")
def built_lib():
    """Builds libsimc_b200.so in-tree if it is missing (nvcc cross-compiles without a GPU)."""
    from simc_gfortran_b200 import lib_path
    if not os.path.exists(lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return lib_path()
