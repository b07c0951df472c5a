import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PKG_NAME = "mixed-precision_lu_factorization_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mplu():
    """The product package (ctypes binding of libmplu.so).  Builds the library if it is missing."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build(with_reference=False)
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def oracle():
    import mplu_oracle
    return mplu_oracle


@pytest.fixture(scope="session")
def solver(mplu):
    s = mplu.Solver(0)
    yield s
    s.close()
