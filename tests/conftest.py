import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def nfn_lib():
    """libnfn_b200.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    from normalizingflownetwork_b200 import _lib
    from normalizingflownetwork_b200 import build as nfn_build

    if not os.path.exists(_lib.LIB_PATH):
        nfn_build.build(verbose=False)
    return _lib.load()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test is marked gpu and needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda:0")
