import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The CUDA library must exist for every test (the CPU tests check that it loads and exports the ABI)."""
    from l3embedding_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()
