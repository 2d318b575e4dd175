"""Host-side parameter selection around the device criterion: the caller of the hot path.

GPmp drives the inner loop from `gpmp/kernel/parameter_selection.py` (SciPy SLSQP / L-BFGS-B over a criterion
closure, :35-437, :583-1577) and seeds it with `gpmp/kernel/init.py:24-66`.  Those modules are plain host control
flow and can be used unchanged on top of `gpmp_b200.num` / `gpmp_b200.core`; this file restates the small part of
them that the BASELINE configs exercise (array data, ML / REML, optional parameterised mean) so that the
package is usable on its own: same function names, argument order, defaults and the same info fields.
Dataloader / mini-batch selection, priors (REMAP) and custom bounds helpers are not restated.
"""
from __future__ import annotations

import math
import time

import numpy as np
from scipy.optimize import OptimizeResult, minimize

from . import num, ops


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


# ---- criterion closures (kernel/parameter_selection.py:35-124) --------------------------------------------------
def make_selection_criterion_with_gradient(model, selection_criterion, xi, zi, parameterized_mean=False,
                                           meanparam_len=1):
    if parameterized_mean:
        def crit_(param, x, z):
            return selection_criterion(model, param[:meanparam_len], param[meanparam_len:], x, z)
    else:
        def crit_(covparam, x, z):
            return selection_criterion(model, covparam, x, z)
    crit = num.DifferentiableSelectionCriterion(crit_, xi, zi)
    return crit.evaluate, crit.evaluate_pre_grad, crit.evaluate_no_grad, crit.gradient


# ---- optimiser driver (kernel/parameter_selection.py:128-276) ----------------------------------------------------
def autoselect_parameters(p0, criterion, gradient, bounds=None, bounds_auto=True, bounds_delta=10.0, silent=True,
                          info=False, method="SLSQP", method_options=None):
    tic = time.time()
    p0 = np.asarray(num.to_np(p0), dtype=np.float64)
    if bounds is None and bounds_auto:
        bounds = [(max(v - bounds_delta, -500.0), min(v + bounds_delta, 500.0)) for v in p0]
    hist_p, hist_j = [], []
    best = {"J": float("inf"), "p": None}

    def fun(p):
        try:
            J = criterion(p)
        except Exception as exc:  # noqa: BLE001 - linear-algebra failures count as +inf, like the reference
            if num._is_linalg_exception(exc):
                J = np.inf
            else:
                raise
        J = float(J)
        hist_p.append(p.copy())
        hist_j.append(J)
        if J < best["J"]:
            best["J"], best["p"] = J, p.copy()
        return J

    def jac(p):
        return np.asarray(num.to_np(gradient(p)), dtype=np.float64)

    options = {"disp": not silent}
    if method == "L-BFGS-B":
        # the reference also passes disp / iprint=-1 (parameter_selection.py:236-247); current SciPy has dropped
        # both for this solver and only warns about them, so they are left out
        options = dict(maxcor=20, ftol=1e-6, gtol=1e-5, eps=1e-8, maxfun=15000, maxiter=15000, maxls=40)
    elif method == "SLSQP":
        options.update(dict(ftol=1e-6, eps=1e-8, maxiter=15000))
    else:
        raise ValueError("Optimization method not implemented.")
    options.update(method_options or {})
    r = minimize(fun, p0, method=method, jac=jac, bounds=bounds, options=options)
    best_p = best["p"] if best["p"] is not None else r.x
    if info:
        r = OptimizeResult(r)
        r["history_params"] = hist_p
        r["history_criterion"] = hist_j
        r["initial_params"] = p0
        r["final_params"] = best_p
        r["bounds"] = bounds
        r["selection_criterion"] = criterion
        r["total_time"] = time.time() - tic
        return best_p, r
    return best_p, None


# ---- front-ends (kernel/parameter_selection.py:280-437, 583-800) ---------------------------------------------------
def select_parameters_with_criterion(model, criterion, xi=None, zi=None, meanparam0=None, covparam0=None,
                                     parameterized_mean=False, meanparam_len=1, info=False, verbosity=0, *,
                                     bounds=None, bounds_auto=True, bounds_delta=10.0, method="SLSQP",
                                     method_options=None):
    tic = time.time()
    if covparam0 is None:
        covparam0 = anisotropic_parameters_initial_guess(model, xi, zi)
    covparam0 = np.asarray(num.to_np(covparam0), dtype=np.float64)
    if parameterized_mean:
        if meanparam0 is None:
            raise ValueError("meanparam0 must be provided when parameterized_mean=True.")
        param0 = np.concatenate((np.asarray(num.to_np(meanparam0), dtype=np.float64).reshape(-1), covparam0))
    else:
        param0 = covparam0
    crit, crit_pre_grad, crit_no_grad, crit_grad = make_selection_criterion_with_gradient(
        model, criterion, xi, zi, parameterized_mean=parameterized_mean, meanparam_len=meanparam_len)
    if verbosity == 1:
        print("Parameter selection using custom criterion...")
    param_opt, info_ret = autoselect_parameters(param0, crit_pre_grad, crit_grad, bounds=bounds,
                                                bounds_auto=bounds_auto, bounds_delta=bounds_delta,
                                                silent=verbosity != 2, info=True, method=method,
                                                method_options=method_options)
    if verbosity == 1:
        print("done.")
    if parameterized_mean:
        meanparam_opt, covparam_opt = param_opt[:meanparam_len], param_opt[meanparam_len:]
        model.meanparam = num.asparam(meanparam_opt)
    else:
        meanparam_opt, covparam_opt = None, param_opt
    model.covparam = num.asparam(covparam_opt)
    if info:
        info_ret["meanparam0"] = meanparam0 if parameterized_mean else None
        info_ret["covparam0"] = covparam0
        info_ret["meanparam"] = meanparam_opt
        info_ret["covparam"] = covparam_opt
        info_ret["selection_criterion"] = crit
        info_ret["selection_criterion_nograd"] = crit_no_grad
        info_ret["time"] = time.time() - tic
        return model, info_ret
    return model, None


def select_parameters_with_reml(model, xi=None, zi=None, covparam0=None, info=False, verbosity=0, **kwargs):
    """REML selection of the covariance parameters (kernel/parameter_selection.py:730-800)."""
    return select_parameters_with_criterion(model, negative_log_restricted_likelihood, xi=xi, zi=zi,
                                            covparam0=covparam0, info=info, verbosity=verbosity, **kwargs)


def select_parameters_with_ml_zero_mean(model, xi=None, zi=None, covparam0=None, info=False, verbosity=0, **kwargs):
    """Maximum likelihood for a zero-mean model."""
    if covparam0 is None:
        covparam0 = anisotropic_parameters_initial_guess_zero_mean(model, xi, zi)
    return select_parameters_with_criterion(model, negative_log_likelihood_zero_mean, xi=xi, zi=zi,
                                            covparam0=covparam0, info=info, verbosity=verbosity, **kwargs)


def select_parameters_with_ml(model, xi=None, zi=None, meanparam0=None, covparam0=None, info=False, verbosity=0,
                              **kwargs):
    """Maximum likelihood with a parameterised mean: the optimisation vector is [meanparam, covparam]."""
    if covparam0 is None or meanparam0 is None:
        m0, c0 = anisotropic_parameters_initial_guess_constant_mean(model, xi, zi)
        meanparam0 = m0 if meanparam0 is None else meanparam0
        covparam0 = c0 if covparam0 is None else covparam0
    meanparam0 = np.asarray(num.to_np(meanparam0), dtype=np.float64).reshape(-1)
    return select_parameters_with_criterion(model, negative_log_likelihood, xi=xi, zi=zi, meanparam0=meanparam0,
                                            covparam0=covparam0, parameterized_mean=True,
                                            meanparam_len=len(meanparam0), info=info, verbosity=verbosity, **kwargs)
