"""Seeded inputs shared by the golden-vector generator, the oracle tests and the GPU parity tests.

TEST INFRASTRUCTURE ONLY (same rule as oracle/gp_numpy.py).

Inputs follow SURVEY.md §8(d): numpy.random.default_rng(seed), x ~ U[0,1]^{n x d},
z = sin(3 sum_j x_j) + 0.1 N(0,1).  Nothing here uses gpmp.misc.designs (unseeded).
"""
from __future__ import annotations

import numpy as np

# --- covariance cases: (name, n, m, d, p, isotropic, seed)
COV_CASES = [
    ("cov_d1_p0", 33, 21, 1, 0, False, 11),
    ("cov_d1_p3", 40, 17, 1, 3, False, 12),
    ("cov_d3_p1", 37, 23, 3, 1, False, 13),
    ("cov_d3_p2", 64, 64, 3, 2, False, 14),
    ("cov_d8_p2", 130, 70, 8, 2, False, 15),
    ("cov_d8_p4", 50, 29, 8, 4, False, 16),
    ("cov_d5_p10", 45, 31, 5, 10, False, 17),
    ("cov_d4_p2_iso", 48, 20, 4, 2, True, 18),
    ("cov_d10_p2", 257, 129, 10, 2, False, 19),
]

# --- likelihood cases: (name, n, d, p, mean, noise, seed)
#     mean in {"zero","const","linear","param"}; noise=True -> example07 covariance
LIK_CASES = [
    ("lik_n6_d1_p3_const", 6, 1, 3, "const", False, 21),
    ("lik_n64_d2_p1_zero", 64, 2, 1, "zero", False, 22),
    ("lik_n64_d2_p2_const", 64, 2, 2, "const", False, 23),
    ("lik_n200_d3_p2_linear", 200, 3, 2, "linear", False, 24),
    ("lik_n300_d6_p2_const_noisy", 300, 6, 2, "const", True, 25),
    ("lik_n257_d4_p3_const", 257, 4, 3, "const", False, 26),
    ("lik_n150_d3_p0_const", 150, 3, 0, "const", False, 27),
    ("lik_n120_d2_p2_param", 120, 2, 2, "param", False, 28),
    ("lik_n500_d8_p2_const", 500, 8, 2, "const", False, 29),
    ("lik_n140_d5_p4_iso_const", 140, 5, 4, "const_iso", False, 30),
]

# --- predict cases: (name, n, m, d, p, mean, noise, seed)
PRED_CASES = [
    ("pred_n6_m200_d1_p3_const", 6, 200, 1, 3, "const", False, 41),
    ("pred_n100_m57_d2_p2_zero", 100, 57, 2, 2, "zero", False, 42),
    ("pred_n200_m90_d3_p2_linear", 200, 90, 3, 2, "linear", False, 43),
    ("pred_n150_m64_d2_p1_param", 150, 64, 2, 1, "param", False, 44),
    ("pred_n300_m130_d6_p2_const_noisy", 300, 130, 6, 2, "const", True, 45),
]

# --- batched criterion (SMC boundary) cases: (name, n, d, p, N, seed)
BATCH_CASES = [
    ("batch_n64_d2_p2_N16", 64, 2, 2, 16, 51),
    ("batch_n128_d4_p2_N12", 128, 4, 2, 12, 52),
]

MEANPARAM = np.array([0.3, -0.5])


def mean_fn(kind, xp):
    """Return mean(x, meanparam) for the given kind using array namespace xp (numpy-like
    module with ones/hstack, e.g. numpy or gpmp.num)."""
    if kind in ("const", "const_iso"):
        return lambda x, _mp: xp.ones((x.shape[0], 1))
    if kind == "linear":
        return lambda x, _mp: xp.hstack((xp.ones((x.shape[0], 1)), x))
    if kind == "param":
        return lambda x, mp: (mp[0] + mp[1] * x[:, 0]).reshape(-1, 1)
    if kind == "zero":
        return None
    raise ValueError(kind)


def meantype_of(kind):
    return {"zero": "zero", "const": "linear_predictor", "const_iso": "linear_predictor",
            "linear": "linear_predictor", "param": "parameterized"}[kind]


def data(n, d, seed, m=0):
    """x (n,d), z (n,), xt (m,d)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, d))
    z = np.sin(3.0 * x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    xt = rng.uniform(size=(m, d)) if m else None
    return x, z, xt


def theta(d, seed, noise=False, iso=False, rho=None):
    """covparam = [log s2, (log tau2,) loginvrho...] with moderate length-scales so that the
    nugget-only matrices stay reasonably conditioned at these sizes."""
    rng = np.random.default_rng(seed + 1000)
    if rho is None:
        rho = 0.25 * np.sqrt(d)
    lir = -np.log(rho) + 0.3 * rng.standard_normal(1 if iso else d)
    head = [0.4 * rng.standard_normal()]
    if noise:
        head.append(np.log(0.01) + 0.2 * rng.standard_normal())
    return np.concatenate((np.array(head), lir))


def headline(n=8192, d=8, seed=1234):
    """BASELINE config 3: x ~ U[0,1]^{n x d}, z = sin(3 sum x) + 0.1 N, theta0 = [0, -log 0.7 ...]."""
    x, z, _ = data(n, d, seed)
    th0 = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
    return x, z, th0


# ---------------------------------------------------------------------------------------------------------
# Benchmarked-size cases (BASELINE.json configs 1-5): inputs for oracle/make_golden_large.py and the GPU tests
# ---------------------------------------------------------------------------------------------------------
_H6_ALPHA = np.array([1.0, 1.2, 3.0, 3.2])
_H6_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8],
                  [17, 8, 0.05, 10, 0.1, 14]], dtype=np.float64)
_H6_P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]],
                        dtype=np.float64)


def hartmann6(x):
    """The 6-d Hartmann test function (public definition), the response of BASELINE config 2."""
    d2 = ((x[:, None, :] - _H6_P[None, :, :]) ** 2 * _H6_A[None, :, :]).sum(axis=2)
    return -(np.exp(-d2) * _H6_ALPHA[None, :]).sum(axis=1)


def large_cfg2(n=2000, d=6, seed=2002):
    """config 2: x ~ U[0,1]^{2000x6}, z = hartmann6(x) + 0.1 N(0,1); theta = [log s2, log tau2, log 1/rho_1..6]."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, d))
    z = hartmann6(x) + 0.1 * rng.standard_normal(n)
    th = np.concatenate(([np.log(0.2), np.log(0.01)], -np.log(0.6) + 0.2 * rng.standard_normal(d)))
    return x, z, th


def large_cfg4(n=512, d=4, N=256, seed=4004):
    """config 4: x ~ U[0,1]^{512x4}, z = sin(3 sum x) + 0.1 N; particles theta_hat + U(-2, 2)^{N x 5}."""
    x, z, _ = data(n, d, seed)
    th_hat = theta(d, seed)
    TH = th_hat + np.random.default_rng(seed + 3).uniform(-2.0, 2.0, size=(N, d + 1))
    return x, z, TH


def large_cfg5(n=4096, d=10, m=4096, paths=4, seed=5005):
    """config 5 shape at the size where the reference's dense predictor still fits: predict + conditioning."""
    x, z, xt = data(n, d, seed, m)
    th = theta(d, seed)
    ztsim = np.random.default_rng(seed + 7).standard_normal((n + m, paths))
    return x, z, xt, th, ztsim


def large_cfg3_theta(d=8, seed=3003):
    """A second headline-size parameter vector (theta0 + U(-0.25, 0.25)), as bench.py draws them."""
    th0 = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
    return th0 + np.random.default_rng(seed).uniform(-0.25, 0.25, size=th0.shape)


def example02():
    """config 1 (examples/gpmp_example02_1d_interpolation.py shape): the 6 points / 200 test points / p stored with
    the REML golden run of oracle/make_golden.py."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                             "reference_torch.npz"))
    pre = "select_reml_example02/"
    return g[pre + "x"], g[pre + "z"], g[pre + "xt"], int(g[pre + "p"])


def smc_small(n=40, d=2, seed=6006):
    """A small model for a complete tempered SMC run: data and the sampling box theta_hat +- 3."""
    x, z, _ = data(n, d, seed)
    th = theta(d, seed)
    return x, z, [list(th - 3.0), list(th + 3.0)]
