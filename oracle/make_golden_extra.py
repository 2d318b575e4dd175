"""Golden vectors for the rows added in round 2 (posterior covariance, MH multi-chain, multi-start selection),
produced by running the UNMODIFIED reference (GPmp 0.9.37, /root/reference) in this container and committed as
tests/golden/reference_extra.npz.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_extra.py

  numpy worker: Model.kriging_predictor_with_zero_mean / kriging_predictor with return_type=1 (the full posterior
                covariance of gpmp/core/kriging.py:170-199) on two of the predict cases of oracle/cases.py
  torch worker: an adaptive Metropolis-Hastings run (sample_from_selection_criterion_mh, 4 chains) on the REML
                criterion of a small model with both generators seeded (gnp.set_seed + torch.manual_seed: proposals
                come from torch's default generator, torch_backend.py:1007-1025, uniforms from the backend's own);
                REML selections from four different starting points (the restarts of a multi-start selection)
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

POSTCOV_CASES = ["pred_n100_m57_d2_p2_zero", "pred_n200_m90_d3_p2_linear"]


def worker(backend):
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from oracle import cases, vendor_ref
    vendor_ref.ensure_plot_stub()
    import gpmp as gp
    import gpmp.num as gnp

    tonp = lambda a: np.asarray(gnp.to_np(a), dtype=np.float64)
    cov = lambda p: (lambda x, y, cp, pairwise=False: gp.kernel.maternp_covariance(x, y, p, cp, pairwise))
    out = {}
    if backend == "numpy":
        for name in POSTCOV_CASES:
            _, n, m, d, p, kind, noise, seed = next(c for c in cases.PRED_CASES if c[0] == name)
            x, z, xt = cases.data(n, d, seed, m)
            th = cases.theta(d, seed)
            model = gp.core.Model(cases.mean_fn(kind, gnp), cov(p), None, gnp.asarray(th), cases.meantype_of(kind))
            f = model.kriging_predictor_with_zero_mean if kind == "zero" else model.kriging_predictor
            lam, C = f(gnp.asarray(x), gnp.asarray(xt), 1)
            _, v = f(gnp.asarray(x), gnp.asarray(xt), 0)
            out["postcov_" + name] = dict(theta=th, lam=tonp(lam), cov=tonp(C), var=tonp(v))
    else:
        import torch

        x, z, box = cases.smc_small()
        model = gp.core.Model(lambda x_, mp: gnp.ones((x_.shape[0], 1)), cov(2), None, None)
        xg, zg = gnp.asarray(x), gnp.asarray(z)

        def crit(theta):
            with torch.no_grad():
                return model.negative_log_restricted_likelihood(gnp.asarray(theta), xg, zg)

        th = cases.theta(2, 6006)
        starts = th + np.array([[0.0, 0.0, 0.0], [0.5, -0.5, 0.3], [-0.7, 0.4, 0.6], [0.2, 0.9, -0.4]])
        gnp.set_seed(5)
        torch.manual_seed(5)
        samples, mh = gp.mcmc.sample_from_selection_criterion_mh(
            selection_criterion=crit, param_initial_states=starts, n_chains=4, n_steps_total=160, burnin_period=60,
            sampling_box=box, silent=True, plot_chains=False, plot_empirical_distributions=False)
        out["mh_small"] = dict(starts=starts, samples=tonp(samples), accept=tonp(mh.accept))
        # multi-start: the reference's own selection from each start (n = 64 likelihood case)
        name, n, d, p, kind, noise, seed = next(c for c in cases.LIK_CASES if c[0] == "lik_n64_d2_p2_const")
        x, z, _ = cases.data(n, d, seed)
        th0 = cases.theta(d, seed)
        S = th0 + np.random.default_rng(9).uniform(-1.0, 1.0, size=(4, d + 1))
        funs, pars = [], []
        for s in S:
            m = gp.core.Model(lambda x_, mp: gnp.ones((x_.shape[0], 1)), cov(p), None, None)
            m, info = gp.kernel.select_parameters_with_reml(m, x, z, covparam0=s, info=True)
            funs.append(float(info.fun))
            pars.append(tonp(m.covparam))
        out["multistart_n64"] = dict(starts=S, funs=np.array(funs), covparams=np.array(pars))
    flat = {f"{c}/{k}": np.asarray(v) for c, rec in out.items() for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, f"_extra_{backend}.npz"), **flat)
    print(f"[{backend}] done: {sorted(out)}", flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
        return
    merged = {}
    for backend in ("numpy", "torch"):
        env = dict(os.environ, GPMP_BACKEND=backend, OMP_NUM_THREADS="8", GPMP_LOG_LEVEL="WARNING")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", backend], check=True, env=env)
        part = os.path.join(OUT, f"_extra_{backend}.npz")
        with np.load(part) as zf:
            merged.update({k: zf[k] for k in zf.files})
        os.remove(part)
    path = os.path.join(OUT, "reference_extra.npz")
    np.savez_compressed(path, **merged)
    print(f"wrote {path}: {len(merged)} arrays, {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
