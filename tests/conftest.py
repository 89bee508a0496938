import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vvc-mip-gpu_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O  # oracle/oracle.py -- test infrastructure

    O.build()
    return O


@pytest.fixture(scope="session")
def mip():
    import mipb200

    if not os.path.exists(mipb200.LIB_PATH):
        mipb200.build()
    return mipb200
