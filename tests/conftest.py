import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def product_lib():
    """Build (if stale) and return the path of nrd_sample_b200/libnrd_b200.so."""
    from nrd_sample_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def host_library(product_lib):
    from nrd_sample_b200 import nrd_api
    return nrd_api.NrdLibrary(product_lib)


@pytest.fixture(scope="session")
def reference_host_library():
    """The UNMODIFIED reference host library, when oracle/_ref was built (needs /root/reference at build time)."""
    from nrd_sample_b200 import nrd_api
    from oracle import runner
    runner.build()
    if not os.path.exists(runner.REF_LIB_PATH):
        pytest.skip("oracle/_ref/libnrd_ref.so not built (reference tree not mounted)")
    return nrd_api.NrdLibrary(runner.REF_LIB_PATH)
