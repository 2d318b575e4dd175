"""gpmp.kernel surface for the hot path: half-integer Matern kernels and covariances on the device.

Mirrors gpmp/kernel/matern.py:10-141 (same names, argument meaning, `y is x or y is None` dispatch,
nugget 10 sigma2 eps on the same set, pairwise variants).  `param = [log sigma2, log(1/rho_1..d)]`; a single
`log(1/rho)` entry means an isotropic kernel.

Fused path: inside `capture()` (used by gpmp_b200.core.Model around the user's covariance callable) a
same-set, non-pairwise `maternp_covariance` call returns a `LazyMatern` placeholder instead of an n x n
matrix.  If the callable hands it back untouched, the model evaluates the likelihood with the fused
pipeline (K built tile by tile, never differentiated through); any arithmetic on the placeholder
materialises it, and the composable ops take over.
"""
from __future__ import annotations

import threading

import torch

from . import ops

_state = threading.local()


class capture:
    """Context manager enabling lazy same-set covariances (see module docstring)."""

    def __enter__(self):
        self._prev = getattr(_state, "on", False)
        _state.on = True
        return self

    def __exit__(self, *exc):
        _state.on = self._prev
        return False


class LazyMatern:
    """Placeholder for sigma2 k_p(D(x, x)) + nugget I.  Behaves like the materialised tensor under any
    torch function or arithmetic operator."""

    def __init__(self, x, y, p, param):
        self.x, self.y, self.p, self.param = x, y, int(p), param  # y is None: same set (nugget added)
        self._value = None

    def materialize(self):
        if self._value is None:
            self._value = ops.matern_cov(self.x, self.y, self.p, self.param)
        return self._value

    @property
    def shape(self):
        n = self.x.shape[0]
        return torch.Size((n, n if self.y is None else self.y.shape[0]))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        conv = lambda a: a.materialize() if isinstance(a, LazyMatern) else a
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]


def _binary(name):
    def op(self, other):
        other = other.materialize() if isinstance(other, LazyMatern) else other
        return getattr(self.materialize(), name)(other)

    return op


for _n in ("__add__", "__radd__", "__sub__", "__rsub__", "__mul__", "__rmul__", "__truediv__", "__rtruediv__",
           "__matmul__", "__rmatmul__", "__pow__"):
    setattr(LazyMatern, _n, _binary(_n))
LazyMatern.__neg__ = lambda self: -self.materialize()


def materialize(K):
    return K.materialize() if isinstance(K, LazyMatern) else K


def maternp_kernel(p, h):
    """Matern kernel with half-integer regularity nu = p + 1/2 (kernel/matern.py:32-64)."""
    return ops.maternp_kernel(p, h)


def matern32_kernel(h):
    """Matern 3/2 (kernel/matern.py:10-29): (1 + 2 nu^(1/2) h) exp(-2 nu^(1/2) h), nu = 3/2."""
    return ops.maternp_kernel(1, h)


def maternp_covariance_ii_or_tt(x, p, param, pairwise=False):
    """Covariance of the observations (or predictands) at x (kernel/matern.py:67-94)."""
    if pairwise:
        return ops.matern_cov_pairwise(x, None, p, param)
    xd = ops.to_device(x)
    if getattr(_state, "on", False):
        return LazyMatern(xd, None, p, param)
    return ops.matern_cov(xd, None, p, param)


def maternp_covariance_it(x, y, p, param, pairwise=False):
    """Cross-covariance between x and y, no nugget (kernel/matern.py:97-121)."""
    if pairwise:
        return ops.matern_cov_pairwise(x, y, p, param)
    if getattr(_state, "on", False):
        return LazyMatern(ops.to_device(x), ops.to_device(y), p, param)
    return ops.matern_cov(x, y, p, param)


def maternp_covariance(x, y, p, param, pairwise=False):
    """Matern covariance wrapper (kernel/matern.py:124-141): identity of `y` selects the same-set form."""
    if y is x or y is None:
        return maternp_covariance_ii_or_tt(x, p, param, pairwise)
    return maternp_covariance_it(x, y, p, param, pairwise)


# initial guesses live in gpmp.kernel in the reference (kernel/__init__.py); same names here.  The optimiser
# front-ends (select_parameters_with_*) are GPmp's own: bind this library with gpmp_b200.dropin.install().
from .selection import (  # noqa: E402,F401
    anisotropic_parameters_initial_guess,
    anisotropic_parameters_initial_guess_constant_mean,
    anisotropic_parameters_initial_guess_zero_mean,
    multistart_reml,
    negative_log_likelihood,
    negative_log_likelihood_zero_mean,
    negative_log_restricted_likelihood,
)
