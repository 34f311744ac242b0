"""Shared fixtures.  `-m "not gpu"` covers the oracle, the golden vectors, host logic and the C-ABI symbol
table; `-m gpu` tests are the parity tests proper and call the CUDA path through the C ABI."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import tfo
    tfo.build()
    return tfo.Lib("port")


@pytest.fixture(scope="session")
def s1_frames():
    from topfusion_b200 import synth
    return synth.sequence("S1", 8)


@pytest.fixture(scope="session")
def s0_frames():
    from topfusion_b200 import synth
    return synth.sequence("S0", 12)


def has_cuda() -> bool:
    try:
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not has_cuda():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from topfusion_b200 import capi
    capi.lib()
    return capi


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint32)


def same_bits_nan(a, b) -> np.ndarray:
    """per-element equality where any-NaN == any-NaN"""
    a = np.asarray(a); b = np.asarray(b)
    return (a == b) | (np.isnan(a) & np.isnan(b))
