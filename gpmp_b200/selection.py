"""Initial guesses for the covariance parameters (gpmp/kernel/init.py:24-66): each is ONE
hot-path call (a GLS variance at length-scales taken from the data range), so it lives with the device code.

The optimiser loop itself (`autoselect_parameters`, `select_parameters_with_*`, gpmp/kernel/parameter_selection.py)
is host control flow that is out of scope here: GPmp's own driver runs unchanged on top of this library once
`gpmp_b200.dropin.install()` has bound the model / kernel seams (tests/test_dropin_gpu.py runs its REML and REMAP
selections that way).  `multistart_reml` below is the one front-end this package adds: the restarts of a
multi-start selection evaluated as ONE batched value+gradient sweep per optimiser iteration.
"""
from __future__ import annotations

import math

import numpy as np

from . import ops


# ---- criteria (kernel/parameter_selection.py:560-581) --------------------------------------------------------
def negative_log_restricted_likelihood(model, covparam, xi, zi):
    return model.negative_log_restricted_likelihood(covparam, xi, zi)


def negative_log_likelihood_zero_mean(model, covparam, xi, zi):
    return model.negative_log_likelihood_zero_mean(covparam, xi, zi)


def negative_log_likelihood(model, meanparam, covparam, xi, zi):
    return model.negative_log_likelihood(meanparam, covparam, xi, zi)


# ---- initial guesses (kernel/init.py:24-66) --------------------------------------------------------------------
def _rho_guess(xi):
    x = ops.to_device(xi)
    d = x.shape[1]
    delta = (x.max(dim=0).values - x.min(dim=0).values).cpu().numpy()
    return math.exp(math.lgamma(d / 2 + 1) / d) / math.sqrt(math.pi) * delta


def _covparam(sigma2, rho):
    return np.concatenate(([math.log(float(sigma2))], -np.log(rho)))


def anisotropic_parameters_initial_guess_zero_mean(model, xi, zi):
    rho = _rho_guess(xi)
    cp = _covparam(1.0, rho)
    sigma2 = float(model.norm_k_sqrd_with_zero_mean(xi, zi, cp)) / xi.shape[0]
    return _covparam(sigma2, rho)


def anisotropic_parameters_initial_guess_constant_mean(model, xi, zi):
    rho = _rho_guess(xi)
    cp = _covparam(1.0, rho)
    ztkz, kinv1, kinvz = model.k_inverses(xi, zi, cp)
    mean = (kinvz.sum() / kinv1.sum()).reshape(1).cpu().numpy()
    return mean, _covparam(float(ztkz) / xi.shape[0], rho)


def anisotropic_parameters_initial_guess(model, xi, zi):
    """GLS variance at length-scales proportional to the data range (kernel/init.py:54-66)."""
    rho = _rho_guess(xi)
    cp = _covparam(1.0, rho)
    sigma2 = float(model.norm_k_sqrd(xi, zi, cp)) / xi.shape[0]
    return _covparam(sigma2, rho)


# ---- multi-start selection on the batched value+gradient sweep -------------------------------
def multistart_reml(model, xi, zi, p, starts, kind="reml", max_iter=200, gtol=1e-5, ftol=1e-9, group=None):
    """Minimise the REML (or zero-mean ML) criterion from every row of `starts` (R x (1+d)) at once.

    GPmp restarts `autoselect_parameters` (kernel/parameter_selection.py:128-276) one start at a time, each
    iteration one value + gradient on the same (xi, zi).  Here every iteration evaluates all R current points in
    one `BatchedCriterion.value_and_grad` sweep (C-ABI gpmp_criterion_batched_grad; rows shard over the ranks of
    `group`), and each restart advances by its own L-BFGS direction (memory 8) with Armijo backtracking, also
    batched: the trial points of all restarts that still search form the next sweep.  Box: +-10 around each start,
    like the reference's automatic bounds.  Returns (best_theta, best_value, info) with the per-restart end
    points, values, iteration counts and the number of sweeps."""
    from . import batched

    crit = batched.BatchedCriterion(model, xi, zi, p, kind=kind, group=group)
    X = np.array(starts, dtype=np.float64, copy=True)
    if X.ndim == 1:
        X = X.reshape(1, -1)
    R, dim = X.shape
    lo, hi = X - 10.0, X + 10.0
    F, G = crit.value_and_grad(X)
    sweeps = 1
    F = np.where(np.isfinite(F), F, np.inf)
    S = [[] for _ in range(R)]  # (s, y) pairs per restart
    active = np.isfinite(F)
    iters = np.zeros(R, dtype=int)

    def direction(r):
        q = G[r].copy()
        alphas = []
        for s, y in reversed(S[r]):
            a = s.dot(q) / y.dot(s)
            alphas.append(a)
            q -= a * y
        if S[r]:
            s, y = S[r][-1]
            q *= s.dot(y) / y.dot(y)
        else:
            q /= max(1.0, np.linalg.norm(q))
        for (s, y), a in zip(S[r], reversed(alphas)):
            q += (a - y.dot(q) / y.dot(s)) * s
        return -q

    for _ in range(max_iter):
        idx = np.flatnonzero(active)
        if idx.size == 0:
            break
        Dm = np.array([direction(r) for r in idx])
        slope = np.einsum("ij,ij->i", Dm, G[idx])
        bad = slope >= 0.0  # not a descent direction: fall back to steepest descent and forget the memory
        for k in np.flatnonzero(bad):
            S[idx[k]] = []
            Dm[k] = -G[idx[k]] / max(1.0, np.linalg.norm(G[idx[k]]))
        slope = np.einsum("ij,ij->i", Dm, G[idx])
        step = np.ones(idx.size)
        searching = np.ones(idx.size, dtype=bool)
        Xn, Fn, Gn = X[idx].copy(), F[idx].copy(), G[idx].copy()
        for _ls in range(25):
            k = np.flatnonzero(searching)
            if k.size == 0:
                break
            T = np.clip(X[idx[k]] + step[k, None] * Dm[k], lo[idx[k]], hi[idx[k]])
            Ft, Gt = crit.value_and_grad(T)
            sweeps += 1
            ok = np.isfinite(Ft) & (Ft <= F[idx[k]] + 1e-4 * step[k] * slope[k])
            for j in np.flatnonzero(ok):
                Xn[k[j]], Fn[k[j]], Gn[k[j]] = T[j], Ft[j], Gt[j]
            searching[k[ok]] = False
            step[k[~ok]] *= 0.5
        for k, r in enumerate(idx):
            iters[r] += 1
            if searching[k]:  # no acceptable step: this restart has converged as far as it can
                active[r] = False
                continue
            s, y = Xn[k] - X[r], Gn[k] - G[r]
            if s.dot(y) > 1e-12 * np.linalg.norm(s) * np.linalg.norm(y):
                S[r] = (S[r] + [(s, y)])[-8:]
            small = abs(F[r] - Fn[k]) <= ftol * max(1.0, abs(Fn[k])) or np.max(np.abs(Gn[k])) <= gtol
            X[r], F[r], G[r] = Xn[k], Fn[k], Gn[k]
            if small:
                active[r] = False
    best = int(np.argmin(F))
    info = {"thetas": X, "values": F, "iterations": iters, "sweeps": sweeps, "best_index": best}
    return X[best].copy(), float(F[best]), info
