import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        from pandelos_b200 import native
        return native.load().pd_device_count() > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def engine_lib():
    """The product library (nvcc, sm_100a).  GPU tests fail loudly when it is missing or no device is visible."""
    from pandelos_b200 import build, native
    if not os.path.exists(native.ENGINE_LIB):
        build.build_engine()
    lib = native.load(native.ENGINE_LIB)
    assert lib.pd_device_count() > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return lib
