import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native():
    """The in-tree CUDA library; built here if stale (nvcc cross-compiles without a GPU)."""
    from lasgun_b200 import build
    build.build()
    from lasgun_b200 import _native
    _native.lib()
    return _native


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx(native):
    if native.lib().lgb_device_count() <= 0:
        pytest.fail("-m gpu tests need a GPU: lgb_device_count() == 0 and there is no CPU fallback")
    ctx = native.Context(0)
    yield ctx
    ctx.close()
