import os
import sys
import warnings

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", category=FutureWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly (not skip) if it is missing."""
    import torch
    from instant_angelo_b200 import _lib
    assert torch.cuda.is_available(), "GPU test running without a CUDA device"
    lib = _lib.load()
    arch = lib.ia_device_arch()
    assert arch >= 100, f"expected an sm_100 device, got arch {arch}"
    return lib
