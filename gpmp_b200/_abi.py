"""ctypes binding of libgpmp_b200.so (the C-ABI declared in include/gpmp_b200.h).

This is the one place where Python touches the native library.  There is no CPU or PyTorch fallback:
if the shared library is missing or a call is rejected, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpmp_b200.so")

MAX_DIM = 32
MAX_P = 16
MAX_Q = 31
COV_FULL = 0
COV_LOWER = 1
TRI_NONE, TRI_A_UPPER, TRI_A_LOWER, TRI_B_UPPER, TRI_B_LOWER = 0, 1, 2, 3, 4

KC_MATERN, KC_GEMM, KC_POTF2, KC_CONTRACT, KC_SMALL, KC_BATCHED = range(6)


class GpmpError(RuntimeError):
    pass


class CovSpec(C.Structure):
    """struct gpmp_cov_spec (include/gpmp_b200.h)."""

    _fields_ = [
        ("p", C.c_int),
        ("d", C.c_int),
        ("noise", C.c_int),
        ("reserved", C.c_int),
        ("log_sigma2", C.c_double),
        ("log_tau2", C.c_double),
        ("loginvrho", C.c_double * MAX_DIM),
    ]


_vp, _i, _ll, _sz, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_double
_specp = C.POINTER(CovSpec)

# name -> (restype, argtypes); every symbol declared in include/gpmp_b200.h
SIGNATURES = {
    "gpmp_abi_version": (_i, []),
    "gpmp_launch_count": (C.c_ulonglong, []),
    "gpmp_prof_enable": (_i, [_i]),
    "gpmp_prof_read": (_i, [_i, C.POINTER(_d), C.POINTER(C.c_ulonglong), C.POINTER(_d)]),
    "gpmp_scaled_distance": (_i, [C.POINTER(_d), _i, _vp, _i, _vp, _i, _vp, _ll, _vp]),
    "gpmp_scaled_distance_elementwise": (_i, [C.POINTER(_d), _i, _vp, _vp, _i, _vp, _vp]),
    "gpmp_scaled_distance_backward": (_i, [C.POINTER(_d), _i, _vp, _i, _vp, _i, _vp, _ll, _vp, _vp, _sz, _vp]),
    "gpmp_maternp_kernel": (_i, [_i, _vp, _vp, _vp, _ll, _vp]),
    "gpmp_matern_cov": (_i, [_specp, _vp, _i, _vp, _i, _vp, _ll, _i, _vp]),
    "gpmp_matern_cov_pairwise": (_i, [_specp, _vp, _vp, _i, _vp, _vp]),
    "gpmp_contract_workspace_bytes": (_sz, [_i, _i, _i]),
    "gpmp_matern_cov_backward": (_i, [_specp, _vp, _i, _vp, _i, _vp, _ll, _vp, _vp, _sz, _vp]),
    "gpmp_measure_dmma_peak": (_i, [_i, _i, _vp, C.POINTER(C.c_double), _vp]),
    "gpmp_potrf_workspace_bytes": (_sz, [_i, _i]),
    "gpmp_potrf": (_i, [_vp, _i, _i, _ll, _vp, _sz, _vp, _vp]),
    "gpmp_potri": (_i, [_vp, _i, _ll, _vp, _vp, _vp, _vp, _ll, _vp]),
    "gpmp_trsm_rows": (_i, [_vp, _i, _ll, _vp, _vp, _i, _ll, _i, _vp, _vp]),
    "gpmp_gemm_nt": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _d, _d, _i, _i, _vp]),
    "gpmp_transpose": (_i, [_vp, _ll, _vp, _ll, _i, _i, _vp]),
    "gpmp_lik_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "gpmp_lik_value": (_i, [_specp, _vp, _ll, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp, _vp, _vp]),
    "gpmp_lik_grad": (_i, [_specp, _vp, _i, _i, _vp, _sz, _vp, _vp, _vp, _ll, _vp]),
    "gpmp_lik_dist_block": (_i, [_i]),
    "gpmp_lik_dist_prepare": (_i, [_specp, _vp, _ll, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp, _vp]),
    "gpmp_lik_dist_group": (_i, [_i, _i, _vp, _sz, _i, _vp, _vp, _vp]),
    "gpmp_lik_dist_store": (_i, [_i, _i, _vp, _sz, _i, _vp, _vp]),
    "gpmp_lik_dist_update": (_i, [_i, _i, _vp, _sz, _i, _vp, _i, _i, _vp]),
    "gpmp_lik_dist_finish": (_i, [_i, _i, _vp, _sz, _vp, _vp, _vp]),
    "gpmp_lik_ws_offset": (_sz, [_i, _i, _i]),
    "gpmp_lik_ws_ld": (_ll, [_i]),
    "gpmp_lik_dist_tup_rows": (_i, [_i, _i, _vp, _sz, _i, _i, _vp]),
    "gpmp_lik_dist_kinv_rows": (_i, [_i, _i, _vp, _sz, _i, _i, _vp]),
    "gpmp_lik_dist_u_cols": (_i, [_i, _i, _vp, _sz, _i, _i, _vp]),
    "gpmp_lik_dist_contract_rows": (_i, [_specp, _vp, _i, _i, _vp, _sz, _i, _i, _vp, _vp]),
    "gpmp_lik_loo": (_i, [_i, _i, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "gpmp_predict_scratch_bytes": (_sz, [_i, _i, _i]),
    "gpmp_predict_chunk": (_i, [_specp, _vp, _i, _i, _vp, _sz, _vp, _i, _vp, _vp, _vp, _ll, _vp, _sz, _vp, _vp,
                                _i, _vp]),
    "gpmp_lik_trsm_rows": (_i, [_i, _i, _vp, _sz, _vp, _i, _ll, _i, _vp, _vp]),
    "gpmp_criterion_batched_bytes": (_sz, [_i, _i, _i]),
    "gpmp_criterion_batched": (_i, [_specp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _sz, _vp, _vp, _vp]),
    "gpmp_criterion_batched_grad_bytes": (_sz, [_i, _i, _i, _i]),
    "gpmp_criterion_batched_grad": (_i, [_specp, _vp, _i, _vp, _ll, _i, _vp, _ll, _vp, _i, _vp, _sz, _vp, _vp, _vp,
                                         _vp]),
}

_ERR = {-1: "bad argument", -2: "dimension or regularity out of range", -3: "workspace too small",
        -100: "CUDA launch error", -101: "pointer / leading dimension alignment"}

_lib = None


def lib():
    """Load libgpmp_b200.so (built in-tree by __graft_entry__.build() / gpmp_b200/csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpmpError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gpmp_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise GpmpError(f"{what} failed: {_ERR.get(rc, 'error')} (code {rc})")


def require_cuda():
    if not torch.cuda.is_available():
        raise GpmpError("gpmp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def make_spec(p, d, log_sigma2, loginvrho, noise=False, log_tau2=0.0):
    """Host-side gpmp_cov_spec from python floats / sequences."""
    if not (0 <= int(p) <= MAX_P):
        raise GpmpError(f"Matern regularity p={p} outside [0, {MAX_P}]")
    if not (1 <= int(d) <= MAX_DIM):
        raise GpmpError(f"input dimension d={d} outside [1, {MAX_DIM}]")
    s = CovSpec()
    s.p, s.d, s.noise, s.reserved = int(p), int(d), int(bool(noise)), 0
    s.log_sigma2 = float(log_sigma2)
    s.log_tau2 = float(log_tau2)
    lir = [float(v) for v in loginvrho]
    if len(lir) == 1 and d > 1:  # isotropic: scalar loginvrho broadcast (kernel/matern.py accepts both)
        lir = lir * d
    if len(lir) != d:
        raise GpmpError(f"loginvrho has {len(lir)} entries for d={d}")
    for j in range(MAX_DIM):
        s.loginvrho[j] = lir[j] if j < d else 0.0
    return s


def launch_count():
    return int(lib().gpmp_launch_count())


def prof_enable(flag):
    """0 off, 1 per-launch CUDA events, 2 the same with the look-ahead streams serialised (exclusive times)."""
    lib().gpmp_prof_enable(int(flag))


def prof_read(cls):
    ms, n, w = C.c_double(), C.c_ulonglong(), C.c_double()
    check(lib().gpmp_prof_read(cls, C.byref(ms), C.byref(n), C.byref(w)), "gpmp_prof_read")
    return ms.value, int(n.value), w.value
