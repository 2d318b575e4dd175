"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gpmp_b200.h declares, validates arguments before touching the device, and sizes workspaces
consistently.  No compute call is made (there is no GPU in the build container)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from gpmp_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gpmp_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpmp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_abi.SIGNATURES) == names, "ctypes binding and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gpmp_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.gpmp_abi_version() == 1
    assert lib.gpmp_launch_count() == 0


def test_library_is_sm100a_native():
    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_[0-9]+(?!0a)\b", out.replace("sm_100a", ""))


def test_cov_spec_layout_and_checks():
    s = _abi.make_spec(2, 3, 0.5, [0.1, 0.2, 0.3])
    assert (s.p, s.d, s.noise) == (2, 3, 0)
    assert C.sizeof(_abi.CovSpec) == 4 * 4 + 2 * 8 + 32 * 8
    iso = _abi.make_spec(1, 4, 0.0, [0.7])
    assert [iso.loginvrho[j] for j in range(4)] == [0.7] * 4
    with pytest.raises(_abi.GpmpError):
        _abi.make_spec(17, 3, 0.0, [0.0] * 3)
    with pytest.raises(_abi.GpmpError):
        _abi.make_spec(2, 33, 0.0, [0.0] * 33)
    with pytest.raises(_abi.GpmpError):
        _abi.make_spec(2, 3, 0.0, [0.0] * 2)


def test_argument_validation_without_device():
    lib = _abi.lib()
    s = _abi.make_spec(2, 3, 0.0, [0.0] * 3)
    null = C.c_void_p(0)
    assert lib.gpmp_matern_cov(C.byref(s), null, 10, null, 0, null, 10, 0, null) == -1
    assert lib.gpmp_maternp_kernel(99, null, null, null, 0, null) == -2
    assert lib.gpmp_potrf(null, 10, 10, 10, null, 0, null, null) == -1
    assert lib.gpmp_lik_value(C.byref(s), null, 0, null, 0, null, null, 0, null, 0, null, null, null) == -1
    assert lib.gpmp_gemm_nt(null, 2, null, 2, null, 2, 1, 1, 1, 1.0, 0.0, 0, 0, null) == -1
    assert lib.gpmp_criterion_batched(C.byref(s), null, 1, null, 1, null, null, 0, null, 0, null, null, null) == -1
    bad = (C.c_double * 3)(0.0, 0.0, 0.0)
    assert lib.gpmp_scaled_distance(bad, 40, null, 1, null, 1, null, 1, null) == -2


def test_workspace_sizes_are_consistent():
    lib = _abi.lib()
    n, q, d = 8192, 1, 8
    wv = lib.gpmp_lik_workspace_bytes(n, q, d, 0)
    wg = lib.gpmp_lik_workspace_bytes(n, q, d, 1)
    # value: the (n + q + 1) x n work matrix dominates; gradient adds T, T^T and K^-1
    assert 8 * n * n < wv < 8 * n * n * 1.3  # + two panel buffers and the block inverses
    assert wv + 3 * 8 * n * n <= wg < wv + 3 * 8 * n * n * 1.05
    assert lib.gpmp_lik_workspace_bytes(n, 40, d, 0) == 0  # q > GPMP_MAX_Q
    assert lib.gpmp_potrf_workspace_bytes(n, n + 2) > 2 * 8 * n * 512
    b1 = lib.gpmp_criterion_batched_bytes(512, 1, 1)
    b9 = lib.gpmp_criterion_batched_bytes(512, 1, 9)
    assert (b9 - b1) % 8 == 0 and (b9 - b1) // 8 >= 8 * 512 * 514
    assert lib.gpmp_contract_workspace_bytes(8192, 8192, 8) == (128 * 128) * 10 * 8
    assert lib.gpmp_predict_scratch_bytes(8192, 1, 1000) >= 1000 * 512 * 8


def test_missing_cuda_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("only meaningful on the CPU-only build box")
    import gpmp_b200 as gp

    with pytest.raises(_abi.GpmpError):
        gp.num.asarray(np.zeros((3, 2)))
    with pytest.raises(_abi.GpmpError):
        gp.kernel.maternp_covariance(np.zeros((3, 2)), None, 2, np.zeros(3))
