"""CPU oracle (torch-CPU autograd) for likelihood value + covparam gradient.

TEST INFRASTRUCTURE ONLY -- see the header of ``oracle/gp_numpy.py``; the same
import rule applies.

The reference has gradients only under its torch backend, by reverse-mode
autograd through every dense op (gpmp/num/torch_backend.py:547-604,
``gnp.value_and_grad`` :516-533).  This file restates that path with the same
torch calls at the same call sites so that (a) it is the *gradient oracle* and
(b) it is the CPU baseline timed for the headline metric "REML logL+grad
evals/s"  -- the same algorithm (mm-based cdist, complete QR,
two dense GEMMs for W'KW, Cholesky, autograd backward), the same library
(torch CPU LAPACK/BLAS).  Pinned against reference outputs in
``tests/golden/`` by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import math

import torch

_DT = torch.float64
EPS = torch.finfo(_DT).eps
FMAX = torch.finfo(_DT).max


def _custom_sqrt(x):
    """Gradient-safe sqrt at 0 (torch_backend.py:783-788)."""
    mask = x == 0.0
    xc = torch.where(mask, 1.0, x)
    return torch.where(mask, 0.0, torch.sqrt(xc))


def cdist(x, y, zero_diagonal=True):
    """|x|^2 + |y|^2 - 2 x.y expansion (torch_backend.py:791-807)."""
    if x is y:
        xn = (x**2).sum(1).view(-1, 1)
        d2 = xn + xn.t() - 2.0 * torch.mm(x, x.t())
    else:
        xn = (x**2).sum(1).view(-1, 1)
        yn = (y**2).sum(1).view(1, -1)
        d2 = xn + yn - 2.0 * torch.mm(x, y.t())
    d = _custom_sqrt(d2.clamp(min=0.0))
    if zero_diagonal and x is y:
        mask = torch.eye(d.size(0), dtype=torch.bool)
        d = d.masked_fill(mask, 0.0)
    return d


def scaled_distance(loginvrho, x, y):
    """torch_backend.py:810-820."""
    invrho = torch.exp(loginvrho)
    xs = invrho * x
    if x is y:
        return cdist(xs, xs)
    return cdist(xs, invrho * y)


def matern_coefficients(p: int):
    gln = torch.lgamma(torch.arange(2 * p + 2, dtype=_DT))
    return [
        torch.exp(gln[p + 1] - gln[2 * p + 1] + gln[p + i + 1] - gln[i + 1] - gln[p - i + 1])
        for i in range(p)
    ]


def maternp_kernel(p: int, h):
    """kernel/matern.py:54-64 with torch ops."""
    h = torch.where(torch.isinf(h), torch.full_like(h, FMAX / 1000.0), h)
    c = 2.0 * math.sqrt(p + 0.5)
    twoch = 2.0 * c * h
    poly = torch.ones(h.shape, dtype=_DT)
    for i, a in enumerate(matern_coefficients(p)):
        poly = poly + a * (twoch ** (p - i))
    return torch.exp(-c * h) * poly


def matern_cov_ii(x, p, param, noise=False):
    """kernel/matern.py:67-94 (nugget) or examples/gpmp_example07_nd_regression.py:95-111 (noise)."""
    sigma2 = torch.exp(param[0])
    if noise:
        diag = torch.exp(param[1])
        loginvrho = param[2:]
    else:
        diag = 10.0 * sigma2 * EPS
        loginvrho = param[1:]
    D = scaled_distance(loginvrho, x, x)
    return sigma2 * maternp_kernel(p, D) + diag * torch.eye(D.shape[0], dtype=_DT)


def cholesky_solve(A, b):
    """torch_backend.py:879-885."""
    if b.dim() == 1:
        b = b.reshape(-1, 1)
    L = torch.linalg.cholesky(A)
    y = torch.linalg.solve_triangular(L, b, upper=False)
    x = torch.linalg.solve_triangular(L.t(), y, upper=True)
    return x, L


def nll_zero_mean(x, z, p, covparam, noise=False):
    """core/likelihood.py:18-52 under the torch backend."""
    K = matern_cov_ii(x, p, covparam, noise)
    n = K.shape[0]
    try:
        Kinv_z, C = cholesky_solve(K, z)
    except RuntimeError:
        return torch.tensor(float("inf"), requires_grad=True)
    norm2 = torch.einsum("i..., i...", z, Kinv_z)
    ldet = 2.0 * torch.sum(torch.log(torch.diag(C)))
    return (0.5 * (n * math.log(2.0 * math.pi) + ldet + norm2)).reshape(())


def reml(x, z, P, p, covparam, noise=False):
    """core/likelihood.py:92-129 under the torch backend: complete QR (core/linalg.py:69),
    G = W'(K W) (core/linalg.py:88), Cholesky solve."""
    K = matern_cov_ii(x, p, covparam, noise)
    n, q = P.shape
    Q, _ = torch.linalg.qr(P, mode="complete")
    W = Q[:, q:n]
    Wz = torch.matmul(W.T, z)
    G = torch.matmul(W.T, torch.matmul(K, W))
    try:
        GinvWz, C = cholesky_solve(G, Wz)
    except RuntimeError:
        return torch.tensor(float("inf"), requires_grad=True)
    norm2 = torch.einsum("i..., i...", Wz, GinvWz)
    ldet = 2.0 * torch.sum(torch.log(torch.diag(C)))
    return (0.5 * ((n - q) * math.log(2.0 * math.pi) + ldet + norm2)).reshape(())


def value_and_grad(f, theta):
    """gnp.value_and_grad (torch_backend.py:516-533): zero gradient when the value is not finite."""
    with torch.enable_grad():
        t = theta.detach().clone().requires_grad_(True)
        y = f(t)
        if not torch.isfinite(y):
            return y.detach(), torch.zeros_like(t)
        (g,) = torch.autograd.grad(y, t, allow_unused=True)
        if g is None:
            g = torch.zeros_like(t)
    return y.detach(), g.detach()


def reml_value_and_grad(x, z, P, p, covparam, noise=False):
    """One headline 'eval': REML value + d/d covparam, the reference's way (torch-CPU autograd).
    Accepts numpy or torch inputs; returns (float, numpy array)."""
    xt = torch.as_tensor(x, dtype=_DT)
    zt = torch.as_tensor(z, dtype=_DT)
    th = torch.as_tensor(covparam, dtype=_DT)
    if P is None:
        v, g = value_and_grad(lambda t: nll_zero_mean(xt, zt, p, t, noise), th)
    else:
        Pt = torch.as_tensor(P, dtype=_DT)
        v, g = value_and_grad(lambda t: reml(xt, zt, Pt, p, t, noise), th)
    return float(v), g.numpy()
