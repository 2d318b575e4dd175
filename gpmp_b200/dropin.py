"""Bind gpmp_b200 into an installed GPmp (0.9.37) at run time: the seams of INTEGRATION.md section 3 applied as
monkey-patches, so that GPmp's own drivers -- `gp.kernel.select_parameters_with_reml / _remap`
(gpmp/kernel/parameter_selection.py:280-437, 747-1000), `gp.mcmc.sample_from_selection_criterion_smc`
(gpmp/mcmc/param_posterior.py:658-775), the diagnostics -- run unchanged on top of the B200 hot path.

    import gpmp as gp                 # GPMP_BACKEND=torch
    import gpmp_b200.dropin as b200
    b200.install()                    # from here on Model.* / gp.kernel.maternp_covariance run on the GPU

What is replaced (and nothing else: priors, optimiser loop, samplers' control flow, parameter containers and the
CPU `gnp` namespace they are written against stay GPmp's):

  kernel seam   gpmp.kernel.{maternp_kernel, matern32_kernel, maternp_covariance} and the same names in
                gpmp.kernel.matern (gpmp/kernel/matern.py:10-141)            -> gpmp_b200.kernel
  distance seam gnp.scaled_distance / scaled_distance_elementwise (torch_backend.py:810-829), used by
                user-composed covariances                                    -> gpmp_b200.num
  model seam    gpmp.core.Model methods on the path (gpmp/core/model.py:227-683): the three likelihoods, predict,
                loo, norm_k_sqrd*, k_inverses, fisher_information*, sample_paths, conditional_sample_paths*
                                                                             -> gpmp_b200.core.Model
  sampler seam  the per-particle loop of `logpdf_temp` (mcmc/param_posterior.py:752) and the per-chain loop of
                MetropolisHastings.run_samples (mcmc/mh.py:412-443) when the selection criterion is a
                `BatchableCriterion`                                         -> gpmp_b200.batched.BatchedCriterion

Conventions that make the seams fit: covariance parameters, priors and the optimiser's vectors stay HOST tensors
(SciPy needs them there); array data (xi, zi, xt) is moved to the current CUDA device at the Model boundary; the
likelihoods return a 0-d HOST tensor with grad_fn, so `priors.py` terms add around it and `torch.autograd.grad`
returns a host gradient.  The user's `mean(x, meanparam)` callable is written against GPmp's
host `gnp`, so it is evaluated on host copies of the points (Model.host_mean).
"""
from __future__ import annotations

import functools
import importlib

import numpy as np

from . import _abi, batched, core, kernel, num

_saved = {}

_MODEL_METHODS = (
    "negative_log_likelihood_zero_mean", "negative_log_likelihood", "negative_log_restricted_likelihood",
    "predict", "loo", "norm_k_sqrd_with_zero_mean", "norm_k_sqrd", "k_inverses",
    "fisher_information", "fisher_information_cpd",
    "sample_paths", "conditional_sample_paths", "conditional_sample_paths_parameterized_mean",
)
_KERNEL_NAMES = ("maternp_kernel", "matern32_kernel", "maternp_covariance")


def _native(ref_model):
    """The gpmp_b200 Model that stands behind a GPmp Model instance (attributes are read at every call, so
    `model.covparam = ...` assignments by GPmp's drivers are seen)."""
    m = core.Model(ref_model.mean, ref_model.covariance, ref_model.meanparam, ref_model.covparam, ref_model.meantype)
    m.host_mean = True
    return m


def _delegate(name):
    def method(self, *args, **kwargs):
        return getattr(_native(self), name)(*args, **kwargs)

    method.__name__ = name
    method.__doc__ = getattr(core.Model, name).__doc__
    return method


class BatchableCriterion:
    """A selection criterion GPmp's samplers can call one theta at a time (`f(theta) -> scalar`, what
    `info.selection_criterion_nograd` is) that ALSO knows how to evaluate a whole particle set in one sweep.
    Pass it as `selection_criterion=` to `gp.mcmc.sample_from_selection_criterion_smc`: once `install()` has run,
    the sampler's `logpdf_temp` uses `batched(x)` instead of its Python loop over particles."""

    def __init__(self, model, xi, zi, p, kind="reml", group=None, max_bytes=None, extra_term=None):
        native = _native(model) if not isinstance(model, core.Model) else model
        self.crit = batched.BatchedCriterion(native, xi, zi, p, kind=kind, group=group, max_bytes=max_bytes)
        # optional vectorised term added to every value (e.g. a negative log-prior): (N, dim) ndarray -> (N,)
        self.extra_term = extra_term
        self.sweeps = 0
        self.evaluations = 0

    def batched(self, thetas):
        th = np.asarray(thetas.detach().cpu() if hasattr(thetas, "detach") else thetas, dtype=np.float64)
        th = th.reshape(1, -1) if th.ndim == 1 else th
        vals = self.crit(th)
        if self.extra_term is not None:
            vals = vals + np.asarray(self.extra_term(th), dtype=np.float64)
        self.sweeps += 1
        self.evaluations += th.shape[0]
        return vals

    def __call__(self, theta):
        return float(self.batched(theta)[0])


def _batched_logpdf(crit, lower, upper):
    """-J/T with the sampling box, same conventions as mcmc/param_posterior.py:739-759."""
    lo = None if lower is None else np.asarray(lower, dtype=np.float64).reshape(-1)
    hi = None if upper is None else np.asarray(upper, dtype=np.float64).reshape(-1)

    def logpdf(x, temperature):
        import torch

        xs = np.asarray(x.detach().cpu() if hasattr(x, "detach") else x, dtype=np.float64)
        single = xs.ndim == 1
        xs2 = xs.reshape(1, -1) if single else xs
        out = np.full(xs2.shape[0], -np.inf)
        inside = np.ones(xs2.shape[0], dtype=bool) if lo is None else np.all((xs2 >= lo) & (xs2 <= hi), axis=1)
        if inside.any():
            vals = crit.batched(xs2[inside])
            out[inside] = np.where(np.isnan(vals), -np.inf, -vals / float(temperature))
        if single:
            return float(out[0])
        return torch.as_tensor(out)

    return logpdf


def _wrap_run_smc(original):
    @functools.wraps(original)
    def run_smc_sampling(*args, **kwargs):
        fn = kwargs.get("logpdf_parameterized_function")
        crit, lower, upper = _criterion_behind(fn)
        if crit is not None:
            kwargs["logpdf_parameterized_function"] = _batched_logpdf(crit, lower, upper)
        return original(*args, **kwargs)

    return run_smc_sampling


def _criterion_behind(fn):
    """The BatchableCriterion (and box) closed over by param_posterior's `logpdf_temp`, if that is what fn is."""
    try:
        cells = _closure_cells(fn)
        f = _closure_cells(cells.get("_criterion_scalar")).get("f")
    except AttributeError:
        return None, None, None
    if isinstance(f, BatchableCriterion):
        lo, hi = cells.get("lower_b"), cells.get("upper_b")
        to_np = lambda b: None if b is None else np.asarray(b.detach().cpu() if hasattr(b, "detach") else b)
        return f, to_np(lo), to_np(hi)
    return None, None, None


def _closure_cells(fn):
    return dict(zip(fn.__code__.co_freevars, (c.cell_contents for c in fn.__closure__ or ())))


def _mh_target_behind(log_target):
    """(criterion, lower, upper, temperature) closed over by param_posterior._make_log_prob_mh's `log_prob`
    (mcmc/param_posterior.py:229-252) when the criterion is a BatchableCriterion, else None."""
    try:
        cells = _closure_cells(log_target)
    except AttributeError:
        return None
    f = cells.get("selection_criterion")
    if not isinstance(f, BatchableCriterion):
        return None
    to_np = lambda b: None if b is None else np.asarray(b.detach().cpu() if hasattr(b, "detach") else b)
    return f, to_np(cells.get("lower_b")), to_np(cells.get("upper_b")), float(cells.get("temperature", 1.0))


def _wrap_mh_run_samples(original):
    """MetropolisHastings.run_samples (mcmc/mh.py:412-443) walks the chains one by one, one `log_target` call per
    chain and step.  The chains of one step are independent, so with a BatchableCriterion behind `log_target` all
    proposals of a step are evaluated in ONE batched sweep.  Proposals and acceptance uniforms are drawn chain by
    chain in the reference's order, so the random stream -- and therefore the trajectory -- is the reference's."""
    import math

    @functools.wraps(original)
    def run_samples(self, n_steps, show_global_progress=False):
        spec = _mh_target_behind(self.log_target)
        if spec is None or not self.symmetric:
            return original(self, n_steps, show_global_progress)
        import gpmp.num as gnp

        crit, lower, upper, temperature = spec
        logpdf = _batched_logpdf(crit, lower, upper)
        C = self.n_chains
        i0 = self.global_iter + 1
        i1 = i0 + n_steps
        for t in range(i0, i1):
            prev = [self.log_target_values[c, t - 1] for c in range(C)]
            stale = [c for c in range(C) if prev[c] is None or bool(gnp.isnan(gnp.asarray(prev[c])))]
            if stale:  # first step of a run: the current states have no stored log-target yet
                vals = logpdf(gnp.to_np(gnp.stack([self.x[c, t - 1] for c in stale])), temperature)
                for c, v in zip(stale, vals):
                    prev[c] = float(v)
            ys, us = [], []
            for c in range(C):
                ys.append(self.prop_rnd(self.x[c, t - 1], c))
                us.append(max(gnp.to_scalar(gnp.rand()), 1e-300))
            new = logpdf(gnp.to_np(gnp.stack(ys)), temperature)
            for c in range(C):
                lp_x, lp_y = float(prev[c]), float(new[c])
                if math.log(us[c]) < lp_y - lp_x:
                    self.x[c, t], self.accept[c, t], self.log_target_values[c, t] = ys[c], True, lp_y
                else:
                    self.x[c, t], self.accept[c, t], self.log_target_values[c, t] = self.x[c, t - 1], False, lp_x
            self.global_iter += 1
            if show_global_progress and self.global_iter % self.options.progress_interval == 0:
                self._print_progress(self.global_iter, self.global_total, self.start_time)
        return gnp.mean(self.accept[:, i0:i1], axis=1)

    return run_samples


def install(gp=None):
    """Apply the seams to the imported GPmp package (default: `import gpmp`).  Idempotent."""
    if _saved:
        return
    _abi.lib()  # fail loudly before touching anything if the native library is missing
    gp = gp if gp is not None else importlib.import_module("gpmp")
    cfg = importlib.import_module("gpmp.config")
    backend = cfg.get_backend() if hasattr(cfg, "get_backend") else None
    if backend not in (None, "torch"):
        raise _abi.GpmpError(f"gpmp_b200 binds under GPMP_BACKEND=torch; the imported GPmp uses {backend!r}")
    gnp = importlib.import_module("gpmp.num")
    kmod = importlib.import_module("gpmp.kernel")
    matern = importlib.import_module("gpmp.kernel.matern")
    model_mod = importlib.import_module("gpmp.core.model")
    post = importlib.import_module("gpmp.mcmc.param_posterior")

    def patch(obj, name, value):
        _saved[(obj, name)] = getattr(obj, name)
        setattr(obj, name, value)

    for name in _KERNEL_NAMES:
        patch(kmod, name, getattr(kernel, name))
        patch(matern, name, getattr(kernel, name))
    patch(matern, "maternp_covariance_ii_or_tt", kernel.maternp_covariance_ii_or_tt)
    patch(matern, "maternp_covariance_it", kernel.maternp_covariance_it)
    patch(gnp, "scaled_distance", num.scaled_distance)
    patch(gnp, "scaled_distance_elementwise", num.scaled_distance_elementwise)
    for name in _MODEL_METHODS:
        if hasattr(model_mod.Model, name):
            patch(model_mod.Model, name, _delegate(name))
    patch(post, "run_smc_sampling", _wrap_run_smc(post.run_smc_sampling))
    mh = importlib.import_module("gpmp.mcmc.mh")
    patch(mh.MetropolisHastings, "run_samples", _wrap_mh_run_samples(mh.MetropolisHastings.run_samples))


def uninstall():
    """Restore GPmp's own implementations (tests)."""
    for (obj, name), value in _saved.items():
        setattr(obj, name, value)
    _saved.clear()
