"""The library bound into the REAL GPmp package (oracle/_ref: the unmodified reference, vendored by
oracle/vendor_ref.py) exactly as INTEGRATION.md section 3 describes, via gpmp_b200.dropin.install():
GPmp's own drivers then run on top of the B200 hot path.

  * config 1: gp.kernel.select_parameters_with_reml AND select_parameters_with_remap
    (gpmp/kernel/parameter_selection.py:747-800, 867-919; priors of gpmp/kernel/priors.py:467-558 added around the
    custom-op scalar) on example02, against the reference's own CPU runs (golden vectors);
  * config 4 (second half): a complete tempered SMC through gp.mcmc.sample_from_selection_criterion_smc
    (gpmp/mcmc/param_posterior.py:658-775 -> gpmp/mcmc/smc.py:1242) with the particle loop of :752 replaced by
    the batched sweep, against particle statistics of reference CPU runs.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bound():
    """(gp, gnp, b200): the vendored reference with the B200 seams installed."""
    from oracle import vendor_ref

    if not vendor_ref.available():
        pytest.skip("oracle/_ref is not present (run `python -m oracle.vendor_ref` where the reference tree exists)")
    assert torch.cuda.is_available(), "these tests need the B200"
    gp = vendor_ref.import_reference("torch")
    import gpmp.num as gnp
    import gpmp_b200.dropin as b200

    b200.install(gp)
    yield gp, gnp, b200
    b200.uninstall()


def _large(case, backend="torch"):
    z = np.load(os.path.join(GOLDEN_DIR, "reference_large.npz"))
    pre = f"{case}/{backend}/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def _model(gp, gnp, p):
    return gp.core.Model(lambda x, param: gnp.ones((x.shape[0], 1)),
                         lambda x, y, cp, pairwise=False: gp.kernel.maternp_covariance(x, y, p, cp, pairwise),
                         None, None)


def test_reference_reml_selection_runs_on_the_device(bound, golden_t):
    gp, gnp, b200 = bound
    import gpmp_b200 as native

    g = golden_t("select_reml_example02")
    x, z, xt, p = g["x"], g["z"], g["xt"], int(g["p"])
    l0 = native._abi.launch_count()
    model, info = gp.kernel.select_parameters_with_reml(_model(gp, gnp, p), x, z, info=True)
    assert native._abi.launch_count() > l0, "the reference's driver did not reach the CUDA kernels"
    e0 = float(np.max(np.abs(np.asarray(info.covparam0) - g["covparam0"])))
    ef = abs(float(info.fun) - float(g["fun"])) / max(1.0, abs(float(g["fun"])))
    ep = float(np.max(np.abs(gnp.to_np(model.covparam) - g["covparam"])))
    mean, var = model.predict(x, z, xt)
    s2 = float(np.exp(g["covparam"][0]))
    print(f"[parity] config1 REML through GPmp's driver: start {e0:.2e}, optimum {ef:.2e}, |dtheta| {ep:.2e}, "
          f"nit {int(info.nit)} (reference {int(g['nit'])})")
    assert e0 <= 1e-8 * max(1.0, float(np.max(np.abs(g["covparam0"]))))
    assert ef <= 1e-6 and ep <= 1e-3
    # predict at the reference's parameters: isolates the predictor from the optimiser's stopping point
    model.covparam = gnp.asarray(g["covparam"])
    mean, var = model.predict(x, z, xt)
    assert np.max(np.abs(mean - g["mean"])) / max(np.sqrt(s2), np.max(np.abs(g["mean"]))) <= 1e-8
    assert np.max(np.abs(var - g["var"])) / s2 <= 1e-8
    assert callable(info.selection_criterion_nograd)


def test_reference_remap_selection_runs_on_the_device(bound):
    """REMAP = REML + log-priors on log sigma2 and log rho (torch scalars composed AROUND the custom-op value, so
    autograd must flow through both)."""
    gp, gnp, b200 = bound
    g = _large("cfg1_remap")
    x, z, xt, p = cases.example02()
    model, info = gp.kernel.select_parameters_with_remap(_model(gp, gnp, p), x, z, info=True)
    probe = float(info.selection_criterion_nograd(gnp.asarray(g["probe_theta"])))
    eprobe = abs(probe - float(g["probe_value"])) / max(1.0, abs(float(g["probe_value"])))
    ef = abs(float(info.fun) - float(g["fun"])) / max(1.0, abs(float(g["fun"])))
    ep = float(np.max(np.abs(gnp.to_np(model.covparam) - g["covparam"])))
    print(f"[parity] config1 REMAP through GPmp's driver: criterion at a probe point {eprobe:.2e}, optimum {ef:.2e}, "
          f"|dtheta| {ep:.2e}, nit {int(info.nit)} (reference {int(g['nit'])})")
    assert eprobe <= 1e-8
    assert ef <= 1e-6 and ep <= 1e-3
    model.covparam = gnp.asarray(g["covparam"])
    mean, var = model.predict(x, z, xt)
    s2 = float(np.exp(g["covparam"][0]))
    assert np.max(np.abs(mean - g["mean"])) / max(np.sqrt(s2), np.max(np.abs(g["mean"]))) <= 1e-8
    assert np.max(np.abs(var - g["var"])) / s2 <= 1e-8


def test_reference_smc_runs_on_batched_sweeps(bound):
    """A full tempered SMC (ESS tempering, residual resampling, 10 MH moves per stage) by the reference's sampler,
    every likelihood sweep one batched call.  The sampler draws from an unseeded generator
    (gpmp/mcmc/smc.py:129,535), so parity is statistical: the particle mean must agree with the reference's CPU runs
    within the spread of those runs plus Monte-Carlo error."""
    gp, gnp, b200 = bound
    g = _large("smc_small")
    x, z, box = cases.smc_small()
    model = _model(gp, gnp, 2)
    crit = b200.BatchableCriterion(model, x, z, 2, kind="reml")
    particles, smc = gp.mcmc.sample_from_selection_criterion_smc(
        selection_criterion=crit, init_box=box, sampling_box=box, n_particles=400, mh_steps=10)
    P = gnp.to_np(particles)
    assert P.shape == (400, 3)
    assert crit.sweeps > 20 and crit.evaluations >= 400 * crit.sweeps * 0.5, "the per-particle loop was used"
    ref_mean, ref_std = g["means"].mean(axis=0), g["stds"].mean(axis=0)
    spread = g["means"].std(axis=0)
    dev = np.abs(P.mean(axis=0) - ref_mean)
    tol = 4.0 * (spread + ref_std / np.sqrt(400.0)) + 0.05
    print(f"[parity] SMC posterior mean {P.mean(axis=0)} vs reference {ref_mean} (run-to-run spread {spread}); "
          f"std {P.std(axis=0)} vs {ref_std}; {crit.sweeps} sweeps, {crit.evaluations} evaluations")
    assert np.all(dev <= tol), (dev, tol)
    assert np.all(np.abs(P.std(axis=0) - ref_std) <= 0.5 * ref_std + 0.05)
    # one sweep of the sampler's own logpdf equals the reference's scalar criterion particle by particle
    th = P[:5]
    vals = crit.batched(th)
    for i in range(5):
        v = float(model.negative_log_restricted_likelihood(gnp.asarray(th[i]), gnp.asarray(x), gnp.asarray(z)))
        assert abs(vals[i] - v) <= 1e-9 * max(1.0, abs(v))


def test_reference_mh_runs_on_batched_steps(bound):
    """Adaptive Metropolis-Hastings by the reference's sampler (mcmc/mh.py, 4 chains): one batched sweep per step
    instead of one criterion call per chain.  Proposals and uniforms are drawn in the reference's order from the
    same seeded generators, so the whole trajectory must reproduce the reference's CPU run."""
    gp, gnp, b200 = bound
    z_ = np.load(os.path.join(GOLDEN_DIR, "reference_extra.npz"))
    g = {k.split("/", 1)[1]: z_[k] for k in z_.files if k.startswith("mh_small/")}
    x, z, box = cases.smc_small()
    crit = b200.BatchableCriterion(_model(gp, gnp, 2), x, z, 2, kind="reml")
    gnp.set_seed(5)
    torch.manual_seed(5)
    samples, mh = gp.mcmc.sample_from_selection_criterion_mh(
        selection_criterion=crit, param_initial_states=g["starts"], n_chains=4, n_steps_total=160, burnin_period=60,
        sampling_box=box, silent=True, plot_chains=False, plot_empirical_distributions=False)
    S = gnp.to_np(samples)
    same_accepts = float(np.mean(gnp.to_np(mh.accept) == g["accept"]))
    err = float(np.max(np.abs(S - g["samples"])))
    print(f"[parity] MH 4 chains x 160 steps: {crit.sweeps} sweeps for {crit.evaluations} evaluations; identical "
          f"accept/reject decisions {same_accepts:.3f}; max |sample - reference| {err:.2e}")
    assert S.shape == g["samples"].shape
    assert crit.sweeps <= 161 + 1 and crit.evaluations >= 4 * 160
    assert same_accepts == 1.0 and err <= 1e-8
