import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")

# padded sizes outside the compiled menu are compiled on demand by nvcc (lenslesspicam_b200/_jit.py, ~15 s each); the
# suite keeps that off -- such shapes run the cuFFT engine -- except in the test that exercises it
os.environ.setdefault("LPC_JIT", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def _wide(a):
    """float64 / complex128 view of an array (complex values keep their imaginary part)"""
    a = np.asarray(a)
    return a.astype(np.complex128) if np.iscomplexobj(a) else a.astype(np.float64)


def rel_to_max(a, b):
    """max |a-b| / max |b| -- the parity metric (north_star: 1e-5 relative)."""
    a, b = _wide(a), _wide(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rel_l2(a, b):
    a, b = _wide(a), _wide(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


@pytest.fixture(scope="session")
def golden():
    return load_golden
