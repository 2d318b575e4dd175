"""Generate tests/golden/*.npz by running the UNMODIFIED reference (GPmp 0.9.37) from
/root/reference in THIS container.  Run once by the builder and committed; never run on
the GPU box (the reference does not travel).

    python oracle/make_golden.py            # driver: spawns one worker per backend
    GPMP_BACKEND=numpy python oracle/make_golden.py --worker numpy
    GPMP_BACKEND=torch python oracle/make_golden.py --worker torch

The reference fixes its backend at import (gpmp/num/__init__.py:25-34), hence one
subprocess per backend: NumPy/SciPy = value oracle, torch-CPU autograd = gradient oracle.
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


def _load_cases():
    sys.path.insert(0, ROOT)
    from oracle import cases  # noqa

    return cases


def worker(backend):
    sys.path.insert(0, REF)
    cases = _load_cases()
    import gpmp as gp
    import gpmp.num as gnp

    assert gp.config.get_config().backend == backend if hasattr(gp, "config") else True

    def cov_fn(p, noise):
        if not noise:
            def k(x, y, covparam, pairwise=False):
                return gp.kernel.maternp_covariance(x, y, p, covparam, pairwise)
            return k

        # user-composed noisy kernel, examples/gpmp_example07_nd_regression.py:95-130
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

    def model_for(kind, p, noise, covparam):
        mean = cases.mean_fn(kind, gnp)
        mp = gnp.asarray(cases.MEANPARAM) if kind == "param" else None
        return gp.core.Model(mean, cov_fn(p, noise), meanparam=mp, covparam=gnp.asarray(covparam),
                             meantype=cases.meantype_of(kind))

    tonp = lambda a: np.asarray(gnp.to_np(a), dtype=np.float64)
    out = {}

    if backend == "numpy":
        # ---- covariance / distance / kernel
        for name, n, m, d, p, iso, seed in cases.COV_CASES:
            x, _, y = cases.data(n, d, seed, m)
            th = cases.theta(d, seed, iso=iso)
            xg, yg, thg = gnp.asarray(x), gnp.asarray(y), gnp.asarray(th)
            k = min(n, m)
            xk, yk = gnp.asarray(x[:k]), gnp.asarray(y[:k])
            h = np.concatenate(([0.0, 1e-300, 1e-8, np.inf], np.linspace(0.01, 30.0, 60)))
            out[name] = dict(
                x=x, y=y, theta=th, p=p,
                D=tonp(gnp.scaled_distance(thg[1:], xg, yg)),
                Dxx=tonp(gnp.scaled_distance(thg[1:], xg, xg)),
                De=tonp(gnp.scaled_distance_elementwise(thg[1:], xk, yk)),
                Kii=tonp(gp.kernel.maternp_covariance(xg, xg, p, thg)),
                Kii_none=tonp(gp.kernel.maternp_covariance(xg, None, p, thg)),
                Kit=tonp(gp.kernel.maternp_covariance(xg, yg, p, thg)),
                Kii_pw=tonp(gp.kernel.maternp_covariance(xg, None, p, thg, True)),
                Kit_pw=tonp(gp.kernel.maternp_covariance(xk, yk, p, thg, True)),
                h=h, kh=tonp(gp.kernel.maternp_kernel(p, gnp.asarray(h))),
            )
        # ---- likelihood values, RKHS norms, LOO
        for name, n, d, p, kind, noise, seed in cases.LIK_CASES:
            x, z, _ = cases.data(n, d, seed)
            th = cases.theta(d, seed, noise=noise, iso=kind.endswith("_iso"))
            model = model_for(kind, p, noise, th)
            xg, zg, thg = gnp.asarray(x), gnp.asarray(z), gnp.asarray(th)
            rec = dict(x=x, z=z, theta=th, p=p)
            rec["nll_zero"] = float(model.negative_log_likelihood_zero_mean(thg, xg, zg))
            rec["norm_zero"] = float(model.norm_k_sqrd_with_zero_mean(xg, zg, thg))
            zkz, ki1, kiz = model.k_inverses(xg, zg, thg)
            rec["kinv_zkz"], rec["kinv_1"], rec["kinv_z"] = float(zkz), tonp(ki1), tonp(kiz)
            if cases.meantype_of(kind) == "linear_predictor":
                rec["reml"] = float(model.negative_log_restricted_likelihood(thg, xg, zg))
                rec["norm_k"] = float(model.norm_k_sqrd(xg, zg, thg))
            if kind == "param":
                rec["nll_param"] = float(model.negative_log_likelihood(gnp.asarray(cases.MEANPARAM), thg, xg, zg))
            zl, sl, el = model.loo(xg, zg)
            rec["loo_z"], rec["loo_s2"], rec["loo_e"] = tonp(zl), tonp(sl), tonp(el)
            out[name] = rec
        # ---- predict + conditional sample paths
        for name, n, m, d, p, kind, noise, seed in cases.PRED_CASES:
            x, z, xt = cases.data(n, d, seed, m)
            if name.startswith("pred_n6_"):   # BASELINE config 1 shape (example02)
                rng = np.random.default_rng(seed)
                x = np.sort(rng.uniform(-1, 1, size=(n, 1)), axis=0)
                z = gp.misc.testfunctions.twobumps(x)
                xt = np.linspace(-1, 1, m).reshape(-1, 1)
            th = cases.theta(d, seed, noise=noise)
            model = model_for(kind, p, noise, th)
            zpm, zpv, lam = model.predict(x, z, xt, return_lambdas=True)
            rng = np.random.default_rng(seed + 7)
            ztsim = rng.standard_normal((n + m, 5))
            xi_ind, xt_ind = np.arange(n), np.arange(n, n + m)
            if kind == "param":
                zc = model.conditional_sample_paths_parameterized_mean(ztsim, x, xi_ind, z, xt, xt_ind, lam)
            else:
                zc = model.conditional_sample_paths(ztsim, xi_ind, z, xt_ind, lam)
            # deterministic part of sample_paths: chol(K(xt,xt)) @ normals
            Ktt = tonp(model.covariance(gnp.asarray(xt), gnp.asarray(xt), model.covparam))
            out[name] = dict(x=x, z=z, xt=xt, theta=th, p=p, mean=tonp(zpm), var=tonp(zpv),
                             lam=tonp(lam), ztsim=ztsim, cond=tonp(zc), Ktt=Ktt)
        # ---- batched criterion values (the per-particle loop of param_posterior.py:752)
        for name, n, d, p, N, seed in cases.BATCH_CASES:
            x, z, _ = cases.data(n, d, seed)
            th0 = cases.theta(d, seed)
            rng = np.random.default_rng(seed + 3)
            TH = th0 + rng.uniform(-2, 2, size=(N, 1 + d))
            model = model_for("const", p, False, th0)
            xg, zg = gnp.asarray(x), gnp.asarray(z)
            vals = np.array([float(model.negative_log_restricted_likelihood(gnp.asarray(TH[i]), xg, zg))
                             for i in range(N)])
            out[name] = dict(x=x, z=z, TH=TH, p=p, vals=vals)

    if backend == "torch":
        import torch
        for name, n, d, p, kind, noise, seed in cases.LIK_CASES:
            x, z, _ = cases.data(n, d, seed)
            th = cases.theta(d, seed, noise=noise, iso=kind.endswith("_iso"))
            model = model_for(kind, p, noise, th)
            xg, zg, thg = gnp.asarray(x), gnp.asarray(z), gnp.asarray(th)
            rec = {}
            v, g = gnp.value_and_grad(lambda t: model.negative_log_likelihood_zero_mean(t, xg, zg), thg)
            rec["nll_zero"], rec["nll_zero_grad"] = float(v), tonp(g)
            if cases.meantype_of(kind) == "linear_predictor":
                v, g = gnp.value_and_grad(lambda t: model.negative_log_restricted_likelihood(t, xg, zg), thg)
                rec["reml"], rec["reml_grad"] = float(v), tonp(g)
            if kind == "param":
                mp = cases.MEANPARAM
                full = gnp.asarray(np.concatenate((mp, th)))
                v, g = gnp.value_and_grad(
                    lambda t: model.negative_log_likelihood(t[:2], t[2:], xg, zg), full)
                rec["nll_param"], rec["nll_param_grad"] = float(v), tonp(g)
            out[name] = rec
        # ---- BASELINE config 1 end to end: REML selection + predict (example02 shape)
        name, n, m, d, p, kind, noise, seed = cases.PRED_CASES[0]
        rng = np.random.default_rng(seed)
        x = np.sort(rng.uniform(-1, 1, size=(n, 1)), axis=0)
        z = gp.misc.testfunctions.twobumps(x)
        xt = np.linspace(-1, 1, m).reshape(-1, 1)
        model = model_for("const", p, False, cases.theta(d, seed))
        model.covparam = None
        model, info = gp.kernel.select_parameters_with_reml(model, x, z, info=True)
        zpm, zpv = model.predict(x, z, xt)
        out["select_reml_example02"] = dict(
            x=x, z=z, xt=xt, p=p, covparam0=tonp(info.covparam0), covparam=tonp(model.covparam),
            fun=float(info.fun), nit=int(info.nit), mean=tonp(zpm), var=tonp(zpv))

    os.makedirs(OUT, exist_ok=True)
    flat = {}
    for case, rec in out.items():
        for k, v in rec.items():
            flat[f"{case}/{k}"] = np.asarray(v)
    path = os.path.join(OUT, f"reference_{backend}.npz")
    np.savez_compressed(path, **flat)
    print(f"[{backend}] wrote {path}: {len(out)} cases, {os.path.getsize(path)/1e6:.2f} MB")


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--worker":
        worker(sys.argv[2])
        return
    for backend in ("numpy", "torch"):
        env = dict(os.environ, GPMP_BACKEND=backend, OMP_NUM_THREADS="8")
        subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", backend], check=True, env=env)


if __name__ == "__main__":
    main()
