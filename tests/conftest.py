import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand, never used by the product."""
    from oracle import fg_oracle_py as fo
    fo.build()
    return fo


@pytest.fixture(scope="session")
def gpu_lib():
    """The C-ABI library; GPU tests fail loudly (not skip) when it is missing or has no device."""
    from feellgood_b200 import capi
    return capi.lib()
