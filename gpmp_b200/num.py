"""The slice of the `gpmp.num` (gnp) namespace that the exact-GP inner loop uses, on the B200.

Same names and calling conventions as gpmp/num/torch_backend.py (the backend GPmp selects with
GPMP_BACKEND=torch); the hot primitives run through libgpmp_b200.so, everything array-shaped lives on
the current CUDA device as float64, and small parameter vectors stay on the host (they travel by value
into the kernels).  Creation / elementwise helpers that user-written mean and covariance callables
typically need are thin device-placing wrappers over torch.
"""
from __future__ import annotations

import builtins
import math

import numpy as np
import torch

from . import _abi, ops

float64 = torch.float64
pi = math.pi
inf = float("inf")
finfo = torch.finfo
eps = torch.finfo(torch.float64).eps
fmax = torch.finfo(torch.float64).max
LinAlgError = torch.linalg.LinAlgError
is_tensor = torch.is_tensor
ndarray = torch.Tensor


# ---- conversion -------------------------------------------------------------------------------------
def asarray(x, dtype=None):
    """Array data -> float64 tensor on the current CUDA device (tensors already there pass through)."""
    return ops.to_device(x)


array = asarray


def asparam(p):
    """Parameter vector -> float64 HOST tensor (covariance parameters are passed by value to the kernels)."""
    if torch.is_tensor(p):
        return p.to(torch.float64)
    return torch.as_tensor(np.asarray(p, dtype=np.float64))


def to_np(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def to_scalar(x):
    return x.item() if torch.is_tensor(x) else float(x)


def _on_device(factory):
    def make(*args, **kwargs):
        kwargs.setdefault("dtype", torch.float64)
        kwargs.setdefault("device", ops.device())
        return factory(*args, **kwargs)

    make.__name__ = factory.__name__
    return make


ones = _on_device(torch.ones)
zeros = _on_device(torch.zeros)
empty = _on_device(torch.empty)
eye = _on_device(torch.eye)
full = _on_device(torch.full)
arange = _on_device(torch.arange)
linspace = _on_device(torch.linspace)

from torch import (  # noqa: E402  (device follows the operands)
    abs, any, all, concatenate, diag, einsum, exp, hstack, isfinite, isinf, isnan, log, matmul, maximum,
    minimum, ones_like, reshape, sqrt, stack, sum, vstack, where, zeros_like,
)


def inftobigf(a, bigf=fmax / 1000.0):
    """torch_backend.py:500-502 (the CUDA Matern kernels apply the same replacement internally)."""
    return torch.where(torch.isinf(a), torch.full_like(a, bigf), a)


def compute_gammaln(up_to_p):
    """Table of gammaln(0..2p+1) (gpmp/num/shared.py:21-41); the device kernels get their Matern
    coefficients from the host the same way (csrc/matern.cu: matern_coef)."""
    return [math.lgamma(i) if i > 0 else float("inf") for i in range(2 * up_to_p + 2)]


def safe_inf():
    """+inf scalar that autograd accepts (torch_backend.py:124-134)."""
    return torch.tensor(float("inf"), requires_grad=True)


def safe_neginf():
    return torch.tensor(-float("inf"), requires_grad=True)


# ---- L0 primitives ------------------------------------------------------------------------------------
def scaled_distance(loginvrho, x, y):
    """torch_backend.py:810-820 / numpy_backend.py:432-436 (direct differences, exact-zero diagonal when
    `y is x or y is None`)."""
    return ops.scaled_distance(loginvrho, x, y)


def scaled_distance_elementwise(loginvrho, x, y):
    """torch_backend.py:823-829."""
    return ops.scaled_distance_elementwise(loginvrho, x, y)


def cholesky(A):
    """Lower Cholesky factor (torch_backend.py:111); raises torch.linalg.LinAlgError when A is not PD.
    The returned tensor remembers its device-side factorisation for the solve helpers below."""
    fac = ops.potrf(A)
    L = fac.lower()
    L._gpmp_factor = fac
    return L


def cholesky_solve(A, b):
    """(A^-1 b, L) with 1-D b treated as a column (torch_backend.py:879-885).  b rides along the
    factorisation as extra rows (forward solve), then one backward row solve finishes it."""
    A = ops.to_device(A)
    b = ops.to_device(b)
    if b.dim() == 1:
        b = b.reshape(-1, 1)
    n = A.shape[0]
    fac = ops.potrf(A, extra_rows=ops.transpose(b))
    rows = fac.A[n:, :]
    ops.trsm_rows(fac, rows, trans=1)
    L = fac.lower()
    L._gpmp_factor = fac
    return ops.transpose(rows[:, :n]).contiguous(), L


def cho_factor(A, lower=False, overwrite_a=False, check_finite=True):
    """torch_backend.py:868-871: (factor, lower flag)."""
    L = cholesky(A)
    if lower:
        return L, True
    U = L.t()
    U._gpmp_factor = L._gpmp_factor
    return U, False


def cho_solve(L_and_lower, b, overwrite_b=False, check_finite=True):
    """torch_backend.py:873-876: solve with a factor returned by cho_factor / cholesky."""
    L, _lower = L_and_lower
    fac = getattr(L, "_gpmp_factor", None)
    if fac is None:
        raise _abi.GpmpError("cho_solve needs a factor produced by gpmp_b200.num.cho_factor / cholesky")
    b = ops.to_device(b)
    vec = b.dim() == 1
    rows = ops.transpose(b.reshape(-1, 1) if vec else b)
    ops.trsm_rows(fac, rows, trans=0)
    ops.trsm_rows(fac, rows, trans=1)
    x = ops.transpose(rows).contiguous()
    return x.reshape(-1) if vec else x


def cholesky_inv(A):
    """A^-1 for a symmetric positive-definite A (torch_backend.py:888-890) via T = L^-1, A^-1 = T^T T."""
    fac = ops.potrf(A)
    Kinv, _, _ = ops.potri(fac)
    n = fac.n
    lo = torch.tril(Kinv[:, :n])
    return lo + torch.tril(lo, -1).t()


def solve_triangular(A, B, trans=0, lower=False, unit_diagonal=False, overwrite_b=False, check_finite=True):
    """Triangular solve against a factor produced by `cholesky` / `cholesky_solve` / `cho_factor`
    (torch_backend.py:841-865).  trans follows SciPy: 0 solves A X = B, 1 / 'T' solves A^T X = B."""
    fac = getattr(A, "_gpmp_factor", None)
    if fac is None or unit_diagonal:
        raise _abi.GpmpError("solve_triangular is available for Cholesky factors made by gpmp_b200.num only")
    B = ops.to_device(B)
    vec = B.dim() == 1
    rows = ops.transpose(B.reshape(-1, 1) if vec else B)
    t = 1 if trans in (1, "T", "C", 2) else 0
    # an upper factor U = L^T:  U X = B  is the L^T solve, U^T X = B the L solve
    use_lt = (t == 1) if lower else (t == 0)
    ops.trsm_rows(fac, rows, trans=1 if use_lt else 0)
    x = ops.transpose(rows).contiguous()
    return x.reshape(-1) if vec else x


# ---- differentiation helpers: the names GPmp's drivers call (API of torch_backend.py:506-604) -------------------------
# Parameter vectors are HOST leaves here (SciPy hands them over and wants the gradient back on the host); the
# criterion they feed is a custom op whose backward runs on the device.
_LINALG_MARKERS = ("not positive definite", "not positive-definite", "singular", "cholesky", "factorization")


def linalg_failure(exc):
    """True for the failures GPmp's optimiser maps to criterion = +inf (a covariance matrix that is not positive
    definite): torch's LinAlgError, or a RuntimeError whose text names such a failure."""
    if isinstance(exc, torch.linalg.LinAlgError):
        return True
    text = str(exc).lower()
    return isinstance(exc, RuntimeError) and builtins.any(m in text for m in _LINALG_MARKERS)


_is_linalg_exception = linalg_failure  # the reference's private name, kept for code written against it


def _host_leaf(p):
    return asparam(p).detach().clone().requires_grad_(True)


def _scalar_of(y):
    if not torch.is_tensor(y):
        raise TypeError(f"the criterion returned {type(y).__name__}; a 0-d torch tensor is required")
    if y.numel() != 1:
        raise ValueError(f"the criterion returned {y.numel()} values; a scalar is required")
    return y.reshape(())


def value_and_grad(f, x):
    """(f(x), df/dx) as detached host tensors.  A non-finite value comes back with a zero gradient -- the
    convention NUTS / SVGD rely on (torch_backend.py:528-529)."""
    leaf = _host_leaf(x)
    with torch.enable_grad():
        y = _scalar_of(f(leaf))
        g = None
        if bool(torch.isfinite(y)):
            (g,) = torch.autograd.grad(y, leaf, allow_unused=True)
    return y.detach(), (torch.zeros_like(leaf) if g is None else g.detach())


def grad(f):
    """x -> df/dx."""
    return lambda x: value_and_grad(f, x)[1]


class DifferentiableSelectionCriterion:
    """The object `make_selection_criterion_with_gradient` (kernel/parameter_selection.py:116-124) builds around a
    criterion f(p, x, z): SciPy calls `evaluate_pre_grad(p)` for the value (a float) and then `gradient(p)` at the
    same p; samplers call `evaluate_no_grad`.  Same four methods as torch_backend.py:547-604.  The data are
    uploaded once here; p stays a host vector."""

    def __init__(self, f, x, z):
        self.f, self.x, self.z = f, asarray(x), asarray(z)
        self._pending = None  # (host leaf, criterion scalar with its graph) of the last evaluate_pre_grad

    def _run(self, p):
        """Criterion at p; a covariance that is not positive definite counts as +inf."""
        try:
            return self.f(p, self.x, self.z)
        except Exception as exc:  # noqa: BLE001 - everything that is not a linear-algebra failure propagates
            if linalg_failure(exc):
                return None
            raise

    def evaluate(self, p):
        return self.f(p, self.x, self.z)

    __call__ = evaluate

    def evaluate_no_grad(self, p):
        with torch.no_grad():
            v = self._run(asparam(p))
        return inf if v is None else v

    def evaluate_pre_grad(self, p):
        leaf = _host_leaf(p)
        v = self._run(leaf)
        if v is None:
            v = safe_inf()
        self._pending = (leaf, v)
        return v.item()

    def gradient(self, p, retain=False, allow_unused=True):
        if self._pending is None:
            raise ValueError("gradient(p) needs a preceding evaluate_pre_grad(p)")
        leaf, v = self._pending
        if not torch.equal(asparam(p), leaf.detach()):
            raise ValueError("gradient(p) was called at a different p than evaluate_pre_grad(p)")
        if not bool(torch.isfinite(v)) or v.grad_fn is None:
            return torch.zeros_like(leaf)
        (g,) = torch.autograd.grad(v, leaf, retain_graph=retain, allow_unused=allow_unused)
        if g is None:
            raise RuntimeError("the criterion does not depend on p")
        return g
