import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test was selected but no CUDA device is available (there is no CPU fallback)")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "reference_gpu_golden.npz")
    if not os.path.exists(p):
        pytest.fail("tests/golden/reference_gpu_golden.npz missing (tests/golden/make_golden.py)")
    return np.load(p)
