import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()  # builds liboracle.so on first use
    return oracle


@pytest.fixture(scope="session")
def H():
    """The product package with the CUDA library built (nvcc cross-compiles without a GPU)."""
    import hwbloomradixjoin_b200 as h
    from hwbloomradixjoin_b200 import build
    build.build_library()
    return h


@pytest.fixture(scope="session")
def Hgpu(H):
    if H.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    H.set_quiet(True)
    return H
