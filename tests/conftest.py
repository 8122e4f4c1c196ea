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
    from oracle.cpu import Oracle
    return Oracle(threads=min(16, os.cpu_count() or 1))


@pytest.fixture(scope="session")
def reference():
    from oracle.cpu import Reference, reference_available
    if not reference_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def ctx():
    from voroffset_b200 import _lib
    return _lib.Context(0)
