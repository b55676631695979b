"""pytest configuration: the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda():
    """torch, after checking that a CUDA device and the native library exist."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a -m gpu test was run without a CUDA device")
    from astrophotography_b200 import _native
    _native.load()
    return torch


def bits_equal(a, b):
    """Bit-for-bit equality that treats NaN == NaN and +0 == -0 as in np.array_equal."""
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
