import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Case-keyed view over tests/golden/reference_{numpy,torch}.npz (outputs of the real
    reference, written by oracle/make_golden.py)."""

    def __init__(self, backend):
        self._z = np.load(os.path.join(GOLDEN_DIR, f"reference_{backend}.npz"))
        self.cases = sorted({k.split("/")[0] for k in self._z.files})

    def __call__(self, case):
        pre = case + "/"
        return {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}


@pytest.fixture(scope="session")
def golden_np():
    return Golden("numpy")


@pytest.fixture(scope="session")
def golden_t():
    return Golden("torch")


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def relerr_norm(a, b):
    """max |a-b| / max |b| : the scale-relative error used for vectors whose entries cross zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300)) if a.size else 0.0
