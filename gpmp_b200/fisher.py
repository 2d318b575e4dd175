"""Fisher information of the covariance parameters on the device.

Mirrors `gpmp/core/fisher.py:18-155`: I_ij = 0.5 tr(M^-1 dM_i M^-1 dM_j) with M = K (SPD form) or
M = W^T K W in the contrast space of a linear-predictor mean (CPD form), dK_i by the reference's 5-point central
difference of `model.covariance` (`gpmp/num/shared.py:44-55`), so any user-composed covariance works.

Nothing of size (n-q) x n is formed: W M^-1 W^T is the projector Pi = K^-1 - K^-1 P (P^T K^-1 P)^-1 P^T K^-1, which
the gradient pipeline of the library already builds (gpmp_lik_value + gpmp_lik_grad at z = 0 return 0.5 Pi), so
the work is one factorisation + inverse (n^3) and one DMMA GEMM Pi dK_i per parameter (2 n^3 each).
"""
from __future__ import annotations

import torch

from . import num, ops


def _half_projector(model, xi, theta, contrast):
    K = model.covariance(xi, xi, theta)
    n = K.shape[0]
    P = model._basis(xi).contiguous() if contrast else None
    z0 = torch.zeros(n, dtype=torch.float64, device=K.device)
    Kc = K if K.stride(1) == 1 else K.contiguous()
    state, out = ops.lik_value(None, Kc, None, z0, P, True)
    if ops.read_small(out)[6] != 0.0:
        raise RuntimeError("Covariance matrix not invertible; adjust hyperparameters or add jitter.")
    _, _, half_pi = ops.lik_grad(state, want_dz=False, want_dK=True)
    return half_pi


def _information(model, xi, covparam, epsilon, contrast):
    xi = ops.to_device(xi)
    theta = num.asarray(model.covparam if covparam is None else covparam).detach().clone()
    p = theta.shape[0]
    with torch.no_grad():
        half_pi = ops.padded(_half_projector(model, xi, theta, contrast))
        B = []
        for i in range(p):
            def f(v, i=i):
                t = theta.clone()
                t[i] = v
                return model.covariance(xi, xi, t)

            ti, h = float(theta[i]), float(epsilon)
            dK = (-f(ti + 2 * h) + 8.0 * f(ti + h) - 8.0 * f(ti - h) + f(ti - 2 * h)) / (12.0 * h)
            # Pi dK_i  (dK_i symmetric: the NT product with dK_i as the second operand is Pi dK_i)
            B.append(ops.gemm_nt(half_pi, ops.padded(dK), alpha=2.0))
        info = torch.empty((p, p), dtype=torch.float64, device=half_pi.device)
        for i in range(p):
            for j in range(i, p):
                info[i, j] = info[j, i] = 0.5 * (B[i] * B[j].T).sum()
    return info


def fisher_information(model, xi, covparam=None, epsilon: float = 1e-3):
    """SPD form, M = K (core/fisher.py:18-78)."""
    return _information(model, xi, covparam, epsilon, contrast=False)


def fisher_information_cpd(model, xi, covparam=None, epsilon: float = 1e-3):
    """Contrast-space form when the mean is a linear predictor, else the SPD form (core/fisher.py:81-155)."""
    return _information(model, xi, covparam, epsilon, contrast=model.meantype == "linear_predictor")
