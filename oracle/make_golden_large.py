"""Golden vectors at the BENCHMARKED sizes (BASELINE.json configs 1-5), produced by running the UNMODIFIED
reference (GPmp 0.9.37, /root/reference) in this container.  Run once by the builder and committed as
tests/golden/reference_large.npz; the GPU box never reads /root/reference.

    python oracle/make_golden_large.py            # driver: one worker per backend (fixed at import)

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).  Inputs are regenerated from seeds by
oracle/cases.py (`large_*` helpers), so only outputs are stored (a few hundred KB).

  numpy worker (value oracle: SciPy cdist, LAPACK):
    cfg2_*   n=2000, d=6, noisy composed kernel (examples/gpmp_example07_nd_regression.py:95-130): REML value
    cfg4_*   n=512, d=4: REML values of 256 particles (the loop of mcmc/param_posterior.py:752)
    cfg5_*   n=4096, d=10, nt=4096: universal-kriging predict (mean, var) + 4 conditioned sample paths
    cfg3_*   n=8192, d=8: REML value at the headline theta0
  torch worker (gradient oracle: autograd of the torch backend; also the drivers):
    cfg2_*   REML gradient at theta; one select_parameters_with_criterion run (SLSQP) with #evaluations
    cfg3_*   REML value + gradient at n=8192
    cfg1_*   example02: select_parameters_with_remap (priors around the criterion) + predict
    smc_*    a full tempered SMC run (sample_from_selection_criterion_smc) on a small model: particle mean / std
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def worker(backend):
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from oracle import cases, vendor_ref
    vendor_ref.ensure_plot_stub()
    import gpmp as gp
    import gpmp.num as gnp

    tonp = lambda a: np.asarray(gnp.to_np(a), dtype=np.float64)
    out = {}

    def matern_cov(p):
        return lambda x, y, cp, pairwise=False: gp.kernel.maternp_covariance(x, y, p, cp, pairwise)

    def noisy_cov(p):
        # examples/gpmp_example07_nd_regression.py:95-130
        def k(x, y, param, pairwise=False):
            sigma2 = gnp.exp(param[0])
            loginvrho = param[2:]
            if y is x or y is None:
                if pairwise:
                    return sigma2 * gnp.ones((x.shape[0],))
                K = gnp.scaled_distance(loginvrho, x, x)
                return sigma2 * gp.kernel.maternp_kernel(p, K) + gnp.exp(param[1]) * gnp.eye(K.shape[0])
            if pairwise:
                K = gnp.scaled_distance_elementwise(loginvrho, x, y)
            else:
                K = gnp.scaled_distance(loginvrho, x, y)
            return sigma2 * gp.kernel.maternp_kernel(p, K)
        return k

    const_mean = lambda x, _mp: gnp.ones((x.shape[0], 1))

    # ------------------------------------------------------------------ config 2
    x, z, th = cases.large_cfg2()
    model = gp.core.Model(const_mean, noisy_cov(2), None, gnp.asarray(th))
    xg, zg, thg = gnp.asarray(x), gnp.asarray(z), gnp.asarray(th)
    t0 = time.time()
    if backend == "numpy":
        out["cfg2"] = dict(theta=th, reml=float(model.negative_log_restricted_likelihood(thg, xg, zg)))
    else:
        v, g = gnp.value_and_grad(lambda t: model.negative_log_restricted_likelihood(t, xg, zg), thg)
        rec = dict(theta=th, reml=float(v), reml_grad=tonp(g))
        nev = {"value": 0, "grad": 0}
        crit = gp.kernel.negative_log_restricted_likelihood
        m2, info = gp.kernel.select_parameters_with_criterion(model, crit, xi=x, zi=z, covparam0=th, info=True)
        rec.update(sel_covparam=tonp(m2.covparam), sel_fun=float(info.fun), sel_nit=int(info.nit),
                   sel_nfev=int(info.nfev), sel_njev=int(info.njev))
        out["cfg2"] = rec
    print(f"[{backend}] cfg2 {time.time() - t0:.1f}s", flush=True)

    # ------------------------------------------------------------------ config 4 (values only: numpy)
    if backend == "numpy":
        t0 = time.time()
        x, z, TH = cases.large_cfg4()
        model = gp.core.Model(const_mean, matern_cov(2), None, gnp.asarray(TH[0]))
        xg, zg = gnp.asarray(x), gnp.asarray(z)
        vals = []
        for i in range(TH.shape[0]):
            try:
                vals.append(float(model.negative_log_restricted_likelihood(gnp.asarray(TH[i]), xg, zg)))
            except Exception:  # noqa: BLE001 - the sampler maps linear-algebra failures to +inf
                vals.append(np.inf)
        out["cfg4"] = dict(vals=np.array(vals))
        print(f"[{backend}] cfg4 {time.time() - t0:.1f}s", flush=True)

    # ------------------------------------------------------------------ config 5 shape (numpy)
    if backend == "numpy":
        t0 = time.time()
        x, z, xt, th, ztsim = cases.large_cfg5()
        n, m = x.shape[0], xt.shape[0]
        model = gp.core.Model(const_mean, matern_cov(2), None, gnp.asarray(th))
        zpm, zpv, lam = model.predict(x, z, xt, return_lambdas=True)
        zc = model.conditional_sample_paths(ztsim, np.arange(n), z, np.arange(n, n + m), lam)
        out["cfg5"] = dict(theta=th, mean=tonp(zpm), var=tonp(zpv), cond=tonp(zc),
                           reml=float(model.negative_log_restricted_likelihood(gnp.asarray(th), gnp.asarray(x),
                                                                               gnp.asarray(z))))
        print(f"[{backend}] cfg5 {time.time() - t0:.1f}s", flush=True)

    # ------------------------------------------------------------------ config 3 (headline size)
    t0 = time.time()
    x, z, th0 = cases.headline()
    model = gp.core.Model(const_mean, matern_cov(2), None, gnp.asarray(th0))
    xg, zg = gnp.asarray(x), gnp.asarray(z)
    th1 = cases.large_cfg3_theta()
    if backend == "numpy":
        out["cfg3"] = dict(theta0=th0, reml0=float(model.negative_log_restricted_likelihood(gnp.asarray(th0), xg, zg)))
    else:
        rec = dict(theta0=th0, theta1=th1)
        for tag, th in (("0", th0), ("1", th1)):
            v, g = gnp.value_and_grad(lambda t: model.negative_log_restricted_likelihood(t, xg, zg), gnp.asarray(th))
            rec["reml" + tag], rec["reml_grad" + tag] = float(v), tonp(g)
        out["cfg3"] = rec
    print(f"[{backend}] cfg3 {time.time() - t0:.1f}s", flush=True)

    # ------------------------------------------------------------------ config 1: REMAP selection (torch)
    if backend == "torch":
        t0 = time.time()
        x, z, xt, p = cases.example02()
        model = gp.core.Model(const_mean, matern_cov(p), None, None)
        model, info = gp.kernel.select_parameters_with_remap(model, x, z, info=True)
        zpm, zpv = model.predict(x, z, xt)
        out["cfg1_remap"] = dict(covparam0=tonp(info.covparam0), covparam=tonp(model.covparam), fun=float(info.fun),
                                 nit=int(info.nit), mean=tonp(zpm), var=tonp(zpv))
        # criterion (prior included) at a fixed point, for a driver-independent check
        thp = tonp(info.covparam0) + 0.3
        out["cfg1_remap"]["probe_theta"] = thp
        out["cfg1_remap"]["probe_value"] = float(info.selection_criterion_nograd(gnp.asarray(thp)))
        print(f"[{backend}] cfg1 remap {time.time() - t0:.1f}s", flush=True)

        # -------------------------------------------------------------- tempered SMC on a small model
        t0 = time.time()
        x, z, box = cases.smc_small()
        model = gp.core.Model(const_mean, matern_cov(2), None, None)
        xg, zg = gnp.asarray(x), gnp.asarray(z)

        def crit(theta):
            try:
                with gnp.no_grad() if hasattr(gnp, "no_grad") else _null():
                    return model.negative_log_restricted_likelihood(gnp.asarray(theta), xg, zg)
            except Exception:  # noqa: BLE001
                return gnp.asarray(np.inf)

        means, stds = [], []
        for rep in range(3):
            particles, smc = gp.mcmc.sample_from_selection_criterion_smc(
                selection_criterion=crit, init_box=box, sampling_box=box, n_particles=400, mh_steps=10)
            P = tonp(particles)
            means.append(P.mean(axis=0))
            stds.append(P.std(axis=0))
        out["smc_small"] = dict(box=np.asarray(box), means=np.array(means), stds=np.array(stds))
        print(f"[{backend}] smc {time.time() - t0:.1f}s", flush=True)

    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for case, rec in out.items():
        for k, v in rec.items():
            flat[f"{case}/{backend}/{k}"] = np.asarray(v)
    path = os.path.join(OUT, f"_large_{backend}.npz")
    np.savez_compressed(path, **flat)
    print(f"[{backend}] wrote {path}", flush=True)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
        return
    merged = {}
    for backend in ("numpy", "torch"):
        env = dict(os.environ, GPMP_BACKEND=backend, OMP_NUM_THREADS="8", GPMP_LOG_LEVEL="WARNING")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", backend], check=True, env=env)
        part = os.path.join(OUT, f"_large_{backend}.npz")
        with np.load(part) as zf:
            merged.update({k: zf[k] for k in zf.files})
        os.remove(part)
    path = os.path.join(OUT, "reference_large.npz")
    np.savez_compressed(path, **merged)
    print(f"wrote {path}: {len(merged)} arrays, {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
