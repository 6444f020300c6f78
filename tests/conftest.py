import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_api():
    from checkers import oracle_capi
    return oracle_capi()


@pytest.fixture(scope="session")
def ref_lib():
    from checkers import ref_library
    lib = ref_library()
    if lib is None:
        pytest.skip("oracle/_ref/liblpmp_ref.so not built (reference tree absent at build time)")
    return lib


@pytest.fixture(scope="session")
def cuda_api():
    import lp_mp_b200
    return lp_mp_b200.load_library()
