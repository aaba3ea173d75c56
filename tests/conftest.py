import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) libkp_b200.so; CPU-only nvcc cross-compile."""
    import __graft_entry__ as g
    g.build()
    import kp_b200
    return kp_b200


@pytest.fixture(scope="session")
def cuda_dev(lib_built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
