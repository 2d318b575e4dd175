"""BASELINE config 2 with gpmp_b200: d=6 regression, n=2000 noisy observations, Matern p=2 with a noise variance
composed by the user from the gnp primitives (the covariance of GPmp's examples/gpmp_example07_nd_regression.py:
95-130), REML selection by SciPy's SLSQP with GPmp's options (kernel/parameter_selection.py:236-253) on the device
criterion, prediction on held-out points.  A user-composed covariance takes the composable path: the distance
and Matern ops build K on the device and the likelihood op hands dvalue/dK back to autograd."""
import os
import sys
import time

import numpy as np
import torch
from scipy.optimize import minimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout
import gpmp_b200 as gp

gnp = gp.num
P = 2


def hartmann6(x):
    # the usual 6-D Hartmann function on [0, 1]^6
    a = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    Pm = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                          [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    return -np.sum(a * np.exp(-np.sum(A[None] * (x[:, None, :] - Pm[None]) ** 2, axis=2)), axis=1)


def kernel(x, y, covparam, pairwise=False):
    # covparam = [log sigma2, log noise variance, log 1/rho_1..d]
    sigma2, noise, loginvrho = torch.exp(covparam[0]), torch.exp(covparam[1]), covparam[2:]
    if y is x or y is None:
        if pairwise:
            return sigma2 * gnp.ones((x.shape[0],))
        D = gnp.scaled_distance(loginvrho, x, x)
        return sigma2 * gp.kernel.maternp_kernel(P, D) + noise * gnp.eye(x.shape[0])
    if pairwise:
        return sigma2 * gp.kernel.maternp_kernel(P, gnp.scaled_distance_elementwise(loginvrho, x, y))
    return sigma2 * gp.kernel.maternp_kernel(P, gnp.scaled_distance(loginvrho, x, y))


def main():
    rng = np.random.default_rng(1234)
    n, d, nt = 2000, 6, 1000
    xi = rng.uniform(size=(n, d))
    zi = hartmann6(xi) + 0.1 * rng.standard_normal(n)
    xt = rng.uniform(size=(nt, d))
    model = gp.core.Model(lambda x, meanparam: gnp.ones((x.shape[0], 1)), kernel)
    covparam0 = np.concatenate(([np.log(np.var(zi)), np.log(1e-2)], np.full(d, -np.log(0.5))))
    crit = gnp.DifferentiableSelectionCriterion(
        lambda p_, x_, z_: model.negative_log_restricted_likelihood(p_, x_, z_), xi, zi)
    t0 = time.time()
    r = minimize(crit.evaluate_pre_grad, covparam0, method="SLSQP", jac=lambda p_: np.asarray(crit.gradient(p_)),
                 bounds=[(v - 10.0, v + 10.0) for v in covparam0], options=dict(ftol=1e-6, eps=1e-8, maxiter=15000))
    print(f"REML selection: {r.nfev} evaluations in {time.time() - t0:.2f} s, criterion {float(r.fun):.4f}")
    model.covparam = r.x
    print("covparam:", r.x)
    zpm, zpv = model.predict(xi, zi, xt)
    print("RMSE on held-out points:", float(np.sqrt(np.mean((zpm - hartmann6(xt)) ** 2))))


if __name__ == "__main__":
    main()
