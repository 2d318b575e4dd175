"""CPU oracle (NumPy/SciPy) for the exact-GP inner loop of GPmp 0.9.37.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package ``gpmp_b200`` may
import this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, as the checker
or as the timed CPU baseline, never as the shipped path.

This is a restatement of the reference's *NumPy backend* algorithm for the hot
path (file:line citations are relative to ``/root/reference``).  The arithmetic
itself lives in un-vendored third-party dependencies of the reference
(``numpy>=1.20``, ``scipy>=1.12``, pyproject.toml:25-29; installed here:
numpy 2.3.5 / scipy 1.18.1 + OpenBLAS 0.3.30), so the same library calls are
used at the same call sites: ``scipy.spatial.distance.cdist``
(numpy_backend.py:141,432-436), ``numpy.linalg.cholesky`` +
``scipy.linalg.solve_triangular`` (numpy_backend.py:465-469),
``numpy.linalg.qr(mode="complete")`` (core/linalg.py:69),
``scipy.linalg.solve(assume_a="sym")`` (core/kriging.py:107).

Pinning: the reference's own tests hold no golden vector for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``oracle/make_golden.py`` and
committed under ``tests/golden/`` (``tests/test_oracle_golden.py`` checks
every one of them).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import solve as _solve
from scipy.linalg import solve_triangular as _solve_triangular
from scipy.spatial.distance import cdist as _cdist
from scipy.special import gammaln as _gammaln

EPS = np.finfo(np.float64).eps
FMAX = np.finfo(np.float64).max


# --------------------------------------------------------------------------
# L0: distances (numpy_backend.py:432-447)
# --------------------------------------------------------------------------
def scaled_distance(loginvrho, x, y):
    """D_ik = || exp(loginvrho) * (x_i - y_k) ||_2 via SciPy cdist on pre-scaled points
    (numpy_backend.py:432-436)."""
    invrho = np.exp(np.asarray(loginvrho, dtype=np.float64))
    return _cdist(invrho * x, invrho * y)


def scaled_distance_elementwise(loginvrho, x, y):
    """Row-wise scaled distance; zeros when y is x or None (numpy_backend.py:438-447)."""
    if y is x or y is None:
        return np.zeros((x.shape[0],))
    invrho = np.exp(np.asarray(loginvrho, dtype=np.float64))
    return np.sqrt(np.sum((invrho * (x - y)) ** 2, axis=1))


# --------------------------------------------------------------------------
# L1: Matern kernel / covariance (kernel/matern.py:32-141)
# --------------------------------------------------------------------------
def matern_coefficients(p: int):
    """Coefficients a_i of (2ch)^(p-i), i=0..p-1, evaluated as exp of gammaln differences
    exactly like kernel/matern.py:59-63 (table from num/shared.py:21-41)."""
    gln = _gammaln(np.arange(2 * p + 2))
    return np.array(
        [
            np.exp(gln[p + 1] - gln[2 * p + 1] + gln[p + i + 1] - gln[i + 1] - gln[p - i + 1])
            for i in range(p)
        ],
        dtype=np.float64,
    )


def maternp_kernel(p: int, h):
    """k_p(h) = exp(-c h) (1 + sum_i a_i (2 c h)^(p-i)), c = 2 sqrt(p + 1/2)
    (kernel/matern.py:54-64); inf distances replaced by fmax/1000 (numpy_backend.py:250-252)."""
    h = np.asarray(h, dtype=np.float64)
    h = np.where(np.isinf(h), np.full_like(h, FMAX / 1000.0), h)
    c = 2.0 * math.sqrt(p + 0.5)
    twoch = 2.0 * c * h
    poly = np.ones(h.shape)
    for i, a in enumerate(matern_coefficients(p)):
        poly = poly + a * twoch ** (p - i)
    return np.exp(-c * h) * poly


def maternp_covariance(x, y, p: int, param, pairwise=False):
    """kernel/matern.py:124-141 dispatch on identity; _ii_or_tt :67-94; _it :97-121."""
    param = np.asarray(param, dtype=np.float64)
    sigma2 = np.exp(param[0])
    loginvrho = param[1:]
    if y is x or y is None:
        nugget = 10.0 * sigma2 * EPS
        if pairwise:
            return sigma2 * np.ones((x.shape[0],))
        D = scaled_distance(loginvrho, x, x)
        return sigma2 * maternp_kernel(p, D) + nugget * np.eye(D.shape[0])
    if pairwise:
        D = scaled_distance_elementwise(loginvrho, x, y)
    else:
        D = scaled_distance(loginvrho, x, y)
    return sigma2 * maternp_kernel(p, D)


def matern_noisy_covariance(x, y, p: int, param, pairwise=False):
    """User-composed covariance of examples/gpmp_example07_nd_regression.py:95-130
    (BASELINE config 2): param = [log s2, log tau2, loginvrho...]; K = s2 k_p(D) + tau2 I."""
    param = np.asarray(param, dtype=np.float64)
    sigma2 = np.exp(param[0])
    tau2 = np.exp(param[1])
    loginvrho = param[2:]
    if y is x or y is None:
        if pairwise:
            return sigma2 * np.ones((x.shape[0],))
        D = scaled_distance(loginvrho, x, x)
        return sigma2 * maternp_kernel(p, D) + tau2 * np.eye(D.shape[0])
    if pairwise:
        D = scaled_distance_elementwise(loginvrho, x, y)
    else:
        D = scaled_distance(loginvrho, x, y)
    return sigma2 * maternp_kernel(p, D)


# --------------------------------------------------------------------------
# L0: Cholesky solve (numpy_backend.py:465-469)
# --------------------------------------------------------------------------
def cholesky_solve(A, b):
    L = np.linalg.cholesky(A)
    y = _solve_triangular(L, b, lower=True)
    x = _solve_triangular(L.T, y, lower=False)
    return x, L


# --------------------------------------------------------------------------
# Model: a plain container with the reference's fields (core/model.py:136-166)
# --------------------------------------------------------------------------
class OracleModel:
    def __init__(self, mean, covariance, meanparam=None, covparam=None, meantype="linear_predictor"):
        if meantype not in {"zero", "parameterized", "linear_predictor"}:
            raise ValueError("bad meantype")
        self.mean, self.covariance = mean, covariance
        self.meanparam, self.covparam, self.meantype = meanparam, covparam, meantype


# --------------------------------------------------------------------------
# L2: likelihoods (core/likelihood.py:18-129, core/linalg.py:49-88)
# --------------------------------------------------------------------------
def negative_log_likelihood_zero_mean(model, covparam, xi, zi):
    """core/likelihood.py:18-52."""
    K = model.covariance(xi, xi, covparam)
    n = K.shape[0]
    try:
        Kinv_zi, C = cholesky_solve(K, zi)
    except np.linalg.LinAlgError:
        return np.inf
    norm2 = np.dot(zi, Kinv_zi)
    ldetK = 2.0 * np.sum(np.log(np.diag(C)))
    return 0.5 * (n * np.log(2.0 * np.pi) + ldetK + norm2)


def negative_log_likelihood(model, meanparam, covparam, xi, zi):
    """core/likelihood.py:55-89."""
    zi_prior_mean = np.asarray(model.mean(xi, meanparam)).reshape(-1)
    return negative_log_likelihood_zero_mean(model, covparam, xi, zi - zi_prior_mean)


def contrast_matrix(P):
    """W = Q[:, q:] of the complete QR of P (core/linalg.py:49-70)."""
    n, q = P.shape
    Q, _ = np.linalg.qr(P, mode="complete")
    return Q[:, q:n]


def negative_log_restricted_likelihood(model, covparam, xi, zi):
    """core/likelihood.py:92-129: G = W'(K W), Cholesky, 0.5((n-q)log 2pi + logdet G + quad)."""
    K = model.covariance(xi, xi, covparam)
    P = np.asarray(model.mean(xi, model.meanparam))
    W = contrast_matrix(P)
    Wzi = W.T @ zi
    G = W.T @ (K @ W)
    try:
        GinvWz, C = cholesky_solve(G, Wzi)
    except np.linalg.LinAlgError:
        return np.inf
    norm2 = np.dot(Wzi, GinvWz)
    ldet = 2.0 * np.sum(np.log(np.diag(C)))
    n, q = P.shape
    return 0.5 * ((n - q) * np.log(2.0 * np.pi) + ldet + norm2)


def norm_k_sqrd_with_zero_mean(model, xi, zi, covparam):
    """core/linalg.py:113-118."""
    K = model.covariance(xi, xi, covparam)
    Kinv_zi, _ = cholesky_solve(K, zi)
    return np.dot(zi, Kinv_zi)


def k_inverses(model, xi, zi, covparam):
    """core/linalg.py:121-129 (explicit inverse, numpy_backend.py:458-463)."""
    K = model.covariance(xi, xi, covparam)
    Kinv = np.linalg.inv(K)
    Kinv_zi = Kinv @ zi
    Kinv_1 = Kinv @ np.ones(zi.shape)
    return np.dot(zi, Kinv_zi), Kinv_1, Kinv_zi


def norm_k_sqrd(model, xi, zi, covparam):
    """core/linalg.py:132-141."""
    K = model.covariance(xi, xi, covparam)
    P = np.asarray(model.mean(xi, model.meanparam))
    W = contrast_matrix(P)
    Wzi = W.T @ zi
    G = W.T @ (K @ W)
    GinvWz, _ = cholesky_solve(G, Wzi)
    return np.dot(Wzi, GinvWz)


# --------------------------------------------------------------------------
# Analytic covparam gradient.  The reference obtains the
# gradient by torch autograd (torch_backend.py:547-604); oracle/gp_torch.py restates
# that.  This closed form is a second, independent checker for the CUDA kernels.
# --------------------------------------------------------------------------
def _matern_dk_over_h(p: int, h):
    """k_p'(h)/h = -c^2 exp(-t) q_{p-1}(t) / (2p-1), t = c h (p >= 1); for p = 0, -c exp(-t)/h."""
    c = 2.0 * math.sqrt(p + 0.5)
    t = c * h
    if p == 0:
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.where(h > 0, -c * np.exp(-t) / h, 0.0)
        return out
    # q_{p-1}(t) in the convention k_{p-1}(h') = exp(-t) q_{p-1}(t) with the *same* t
    coef = matern_coefficients(p - 1)
    poly = np.ones_like(t)
    for i, a in enumerate(coef):
        poly = poly + a * (2.0 * t) ** (p - 1 - i)
    return -(c * c) * np.exp(-t) * poly / (2.0 * p - 1.0)


def reml_value_and_grad_analytic(x, z, P, p: int, covparam, noise=False):
    """REML (P is (n,q)) or zero-mean ML (P is None) value and d/d covparam, closed form.
    covparam = [log s2, (log tau2,) loginvrho_1..d]."""
    covparam = np.asarray(covparam, dtype=np.float64)
    n, d = x.shape
    off = 2 if noise else 1
    s2 = np.exp(covparam[0])
    lir = covparam[off:]
    if lir.shape[0] == 1 and d > 1:
        lir = np.repeat(lir, d)
    invrho = np.exp(lir)
    xs = invrho * x
    D = _cdist(xs, xs)
    Kc = s2 * maternp_kernel(p, D)
    diag_add = np.exp(covparam[1]) if noise else 10.0 * s2 * EPS
    K = Kc + diag_add * np.eye(n)
    L = np.linalg.cholesky(K)
    Kinv = np.linalg.inv(K)
    if P is None:
        q = 0
        Pi = Kinv
    else:
        q = P.shape[1]
        KiP = Kinv @ P
        S = P.T @ KiP
        Pi = Kinv - KiP @ np.linalg.solve(S, KiP.T)
    alpha = Pi @ z
    quad = z @ alpha
    ldet = 2.0 * np.sum(np.log(np.diag(L)))
    if P is not None:
        Pt = _solve_triangular(L, P, lower=True)
        _, R1 = np.linalg.qr(Pt)
        _, R0 = np.linalg.qr(P)
        ldet = ldet + 2.0 * np.sum(np.log(np.abs(np.diag(R1)))) - 2.0 * np.sum(np.log(np.abs(np.diag(R0))))
    val = 0.5 * ((n - q) * np.log(2.0 * np.pi) + ldet + quad)
    M = Pi - np.outer(alpha, alpha)
    g = np.zeros_like(covparam)
    g[0] = 0.5 * np.sum(M * Kc) if noise else 0.5 * np.sum(M * K)
    if noise:
        g[1] = 0.5 * np.exp(covparam[1]) * np.trace(M)
    Wk = s2 * _matern_dk_over_h(p, D)
    gl = np.zeros(d)
    for j in range(d):
        dj = (xs[:, j][:, None] - xs[:, j][None, :]) ** 2
        gl[j] = 0.5 * np.sum(M * Wk * dj)
    if covparam[off:].shape[0] == 1 and d > 1:
        g[off] = gl.sum()
    else:
        g[off:] = gl
    return val, g


# --------------------------------------------------------------------------
# L2: kriging predictors (core/kriging.py:35-257) and Model.predict (core/model.py:227-307)
# --------------------------------------------------------------------------
def _posterior_variance(model, xt, lambdamu_t, RHS, return_type=0):
    """core/kriging.py:170-199."""
    if return_type == -1:
        return None
    if return_type == 0:
        prior = model.covariance(xt, None, model.covparam, True)
        return prior - np.einsum("i..., i...", lambdamu_t, RHS)
    if return_type == 1:
        prior = model.covariance(xt, None, model.covparam, False)
        return prior - lambdamu_t.T @ RHS
    raise ValueError("return_type must be in {-1, 0, 1}")


def kriging_predictor_with_zero_mean(model, xi, xt, return_type=0):
    """core/kriging.py:35-67."""
    Kii = model.covariance(xi, xi, model.covparam)
    Kit = model.covariance(xi, xt, model.covparam)
    lambda_t, _ = cholesky_solve(Kii, Kit)
    return lambda_t, _posterior_variance(model, xt, lambda_t, Kit, return_type)


def kriging_predictor(model, xi, xt, return_type=0):
    """core/kriging.py:70-116: saddle-point system solved with scipy solve(assume_a='sym')."""
    Kii = model.covariance(xi, xi, model.covparam)
    Pi = np.asarray(model.mean(xi, model.meanparam))
    ni, q = Pi.shape
    LHS = np.vstack((np.hstack((Kii, Pi)), np.hstack((Pi.T, np.zeros((q, q))))))
    Kit = model.covariance(xi, xt, model.covparam)
    Pt = np.asarray(model.mean(xt, model.meanparam))
    RHS = np.vstack((Kit, Pt.T))
    lambdamu_t = _solve(LHS, RHS, overwrite_a=True, overwrite_b=False, assume_a="sym")
    lambda_t = lambdamu_t[0:ni, :]
    return lambda_t, _posterior_variance(model, xt, lambdamu_t, RHS, return_type)


def predict(model, xi, zi, xt, return_lambdas=False, zero_neg_variances=True):
    """core/model.py:227-307 with core/kriging.py:119-164 (select_predictor)."""
    zi = np.asarray(zi).reshape(-1)
    zt_prior_mean = 0.0
    zi_centered = zi
    if model.meantype == "zero":
        lambda_t, var = kriging_predictor_with_zero_mean(model, xi, xt)
    elif model.meantype == "linear_predictor":
        lambda_t, var = kriging_predictor(model, xi, xt)
    else:
        lambda_t, var = kriging_predictor_with_zero_mean(model, xi, xt)
        zi_centered = zi - np.asarray(model.mean(xi, model.meanparam)).reshape(-1)
        zt_prior_mean = np.asarray(model.mean(xt, model.meanparam)).reshape(-1)
    if zero_neg_variances:
        var = np.maximum(var, 0.0)
    mean = np.einsum("i..., i...", lambda_t, zi_centered) + zt_prior_mean
    if return_lambdas:
        return mean, var, lambda_t
    return mean, var


# --------------------------------------------------------------------------
# L2: LOO (core/loo.py:65-130) -- a SURVEY §8(f) "next" row
# --------------------------------------------------------------------------
def loo(model, xi, zi):
    zi = np.asarray(zi).reshape(-1)
    if model.meantype in ("zero", "parameterized"):
        m = 0.0
        if model.meantype == "parameterized":
            m = np.asarray(model.mean(xi, model.meanparam)).reshape(-1)
        zc = zi - m
        K = model.covariance(xi, xi, model.covparam)
        Kinv_z, C = cholesky_solve(K, zc)
        T = _solve_triangular(C, np.eye(K.shape[0]), lower=True)
        kd = np.sum(T * T, axis=0)
        eloo = Kinv_z / kd
        return zc - eloo + m, 1.0 / kd, eloo
    K = model.covariance(xi, xi, model.covparam)
    P = np.asarray(model.mean(xi, model.meanparam))
    W = contrast_matrix(P)
    G = W.T @ (K @ W)
    S, _ = cholesky_solve(G, W.T)
    Qz = W @ (S @ zi)
    Qd = np.sum(W * S.T, axis=1)
    eloo = Qz / Qd
    return zi - eloo, 1.0 / Qd, eloo


# --------------------------------------------------------------------------
# L2: sample paths (core/sample_paths.py:18-182)
# --------------------------------------------------------------------------
def sample_paths_from_normals(model, xt, normals):
    """Deterministic part of core/sample_paths.py:18-63 ('chol' branch): C @ normals with
    C = chol(K(xt,xt)).  The draw itself (torch_backend.py:912-915) is RNG-specific and is
    not part of parity."""
    K = model.covariance(xt, xt, model.covparam)
    C = np.linalg.cholesky(K)
    return C @ normals


def sample_paths_svd_from_normals(model, xt, normals):
    """Deterministic part of core/sample_paths.py:50-58 ('svd' branch): C @ normals with the symmetric square root
    C = (U sqrt(s)) V^T of K(xt, xt) from gnp.svd(K, hermitian=True) (numpy_backend: numpy.linalg.svd)."""
    K = model.covariance(xt, xt, model.covparam)
    U, s, Vt = np.linalg.svd(K, full_matrices=True, hermitian=True)
    return ((U * np.sqrt(s)) @ Vt) @ normals


def conditional_sample_paths(ztsim, xi_ind, zi, xt_ind, lambda_t):
    """core/sample_paths.py:66-119."""
    zi_ = np.asarray(zi).reshape(-1, 1)
    xi_ind = np.asarray(xi_ind, dtype=int).reshape(-1)
    delta = zi_ - ztsim[xi_ind, :]
    return ztsim[xt_ind, :] + np.einsum("ij,ik->jk", lambda_t, delta)


def conditional_sample_paths_parameterized_mean(model, ztsim, xi, xi_ind, zi, xt, xt_ind, lambda_t):
    """core/sample_paths.py:122-182."""
    xi_ind = np.asarray(xi_ind).reshape(-1)
    xt_ind = np.asarray(xt_ind).reshape(-1)
    zc = np.asarray(zi).reshape(-1) - np.asarray(model.mean(xi, model.meanparam)).reshape(-1)
    mt = np.asarray(model.mean(xt, model.meanparam)).reshape(-1, 1)
    delta = zc.reshape(-1, 1) - ztsim[xi_ind, :]
    return ztsim[xt_ind, :] + np.einsum("ij,ik->jk", lambda_t, delta) + mt


# --------------------------------------------------------------------------
# L4 boundary: particle-batched criterion (mcmc/param_posterior.py:739-759)
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# Fisher information (core/fisher.py:18-155, num/shared.py:44-55)
# --------------------------------------------------------------------------
def _covariance_derivatives(model, xi, theta, epsilon):
    """dK/dtheta_i by the reference's 5-point central difference (num/shared.py:44-55)."""
    out = []
    for i in range(theta.shape[0]):
        def f(v):
            t = theta.copy()
            t[i] = v
            return model.covariance(xi, xi, t)

        h = epsilon
        out.append((-f(theta[i] + 2 * h) + 8 * f(theta[i] + h) - 8 * f(theta[i] - h) + f(theta[i] - 2 * h)) / (12.0 * h))
    return out


def fisher_information(model, xi, covparam=None, epsilon=1e-3):
    """I_ij = 0.5 tr(K^-1 dK_i K^-1 dK_j)  (core/fisher.py:18-78)."""
    theta = np.asarray(model.covparam if covparam is None else covparam, dtype=np.float64)
    Kinv = np.linalg.inv(model.covariance(xi, xi, theta))
    B = [Kinv @ dK for dK in _covariance_derivatives(model, xi, theta, epsilon)]
    p = theta.shape[0]
    return np.array([[0.5 * np.trace(B[i] @ B[j]) for j in range(p)] for i in range(p)])


def fisher_information_cpd(model, xi, covparam=None, epsilon=1e-3):
    """Contrast-space form G = W'KW for a linear-predictor mean (core/fisher.py:81-155)."""
    if model.meantype != "linear_predictor":
        return fisher_information(model, xi, covparam, epsilon)
    theta = np.asarray(model.covparam if covparam is None else covparam, dtype=np.float64)
    W = contrast_matrix(np.asarray(model.mean(xi, model.meanparam)))
    G = W.T @ (model.covariance(xi, xi, theta) @ W)
    B = [np.linalg.solve(G, W.T @ (dK @ W)) for dK in _covariance_derivatives(model, xi, theta, epsilon)]
    p = theta.shape[0]
    return np.array([[0.5 * np.trace(B[i] @ B[j]) for j in range(p)] for i in range(p)])


def logpdf_temp(criterion, x, temperature, lower_b=None, upper_b=None):
    """Serial per-particle loop of param_posterior.py:742-759: -J(theta_i)/T, -inf outside the box."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        if lower_b is not None and (np.any(x < lower_b) or np.any(x > upper_b)):
            return -np.inf
        return -float(criterion(x)) / temperature
    vals = np.asarray([float(criterion(x[i])) for i in range(x.shape[0])])
    out = -vals / temperature
    if lower_b is None:
        return out
    in_box = np.all(x >= lower_b, axis=1) & np.all(x <= upper_b, axis=1)
    return np.where(in_box, out, -np.inf)
