import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gold():
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def crb():
    import cudaraster_linux_b200 as m
    return m


@pytest.fixture(scope="session")
def raster(crb):
    """One CudaRaster context for the GPU tests.  Fails (not skips) when the CUDA extension is missing."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    r = crb.CudaRaster(0)
    yield r
    r.close()
