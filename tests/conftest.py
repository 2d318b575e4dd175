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


# Every error a test measures with relerr / relerr_norm is recorded against the running test, and the largest one
# per test is printed in the terminal summary ("[parity] <test>: ..."), so the margin under each tolerance is visible
# in the log of a passing run.
_CURRENT = {"id": None}
_MEASURED = {}


def _record(v):
    nid = _CURRENT["id"]
    if nid is not None and np.isfinite(v):
        cnt, mx = _MEASURED.get(nid, (0, 0.0))
        _MEASURED[nid] = (cnt + 1, max(mx, v))
    return v


@pytest.fixture(autouse=True)
def _parity_recorder(request):
    _CURRENT["id"] = request.node.nodeid
    yield
    _CURRENT["id"] = None


def pytest_terminal_summary(terminalreporter):
    if not _MEASURED:
        return
    terminalreporter.section("measured parity errors (largest per test)")
    for nid, (cnt, mx) in _MEASURED.items():
        terminalreporter.write_line(f"[parity] {nid}: max of {cnt} measured errors = {mx:.3e}")


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return _record(float(np.max(np.abs(a - b) / den)) if a.size else 0.0)


def relerr_norm(a, b):
    """max |a-b| / max |b| : the scale-relative error used for vectors whose entries cross zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return _record(float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300)) if a.size else 0.0)
