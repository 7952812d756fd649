"""pytest configuration: registers the ``gpu`` marker and shared helpers/fixtures.

``-m "not gpu"`` : oracle vs golden vectors, host logic, C-ABI symbol export, gloo world_size-2 paths.
``-m gpu``       : the parity tests proper -- CUDA path (through the C-ABI) vs the oracle / golden vectors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-for-bit equality, except that any NaN matches any NaN (payload/sign of NaN is not contractual)."""
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    if not np.array_equal(nan_a, nan_b):
        return False
    ai = a.view(np.uint32 if a.dtype == np.float32 else np.uint64 if a.dtype == np.float64 else np.uint16)
    bi = b.view(ai.dtype)
    return bool(np.array_equal(ai[~nan_a], bi[~nan_a]))


def assert_bits_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    if not bits_equal(a, b):
        bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))))
        raise AssertionError(
            f"{what}: {bad.size} of {a.size} elements differ; first idx {bad[:5]}, "
            f"got {a.ravel()[bad[:5]]!r} want {b.ravel()[bad[:5]]!r}")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import adam_oracle as ao
    return ao.COracle()


@pytest.fixture(scope="session")
def np_oracle():
    from oracle import adam_oracle as ao
    return ao.NumpyOracle
