"""Golden vectors for the mini-batch criterion and per-particle value+gradient
(tests/golden/reference_minibatch.npz), produced by the UNMODIFIED reference (torch backend, CPU):
`gnp.BatchDifferentiableSelectionCriterion` (gpmp/num/torch_backend.py:607-718) over a list of batches and
`gnp.value_and_grad` per particle (torch_backend.py:516-533, the SVGD loop of gpmp/mcmc/svgd.py:310-313).

    GPMP_BACKEND=torch python oracle/make_golden_minibatch.py

Test infrastructure only; never run on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

# name, n_total, batch size, d, p, mean kind, seed
MINIBATCH_CASES = [
    ("mb_n500_b64_d3_p2_const", 500, 64, 3, 2, "const", 51),   # 7 batches of 64 + one of 52
    ("mb_n384_b128_d2_p1_zero", 384, 128, 2, 1, "zero", 52),   # 3 equal batches
]
# name, n, d, p, mean kind, number of particles, seed
PARTICLE_CASES = [
    ("vg_n200_d3_p2_const", 200, 3, 2, "const", 6, 53),
    ("vg_n150_d2_p3_zero", 150, 2, 3, "zero", 5, 54),
]


def main():
    os.environ["GPMP_BACKEND"] = "torch"
    import torch
    from oracle import cases
    import gpmp as gp
    import gpmp.num as gnp

    out = {}

    def model_for(kind, p, th):
        return gp.core.Model(cases.mean_fn(kind, gnp),
                             lambda x, y, cp, pairwise=False: gp.kernel.maternp_covariance(x, y, p, cp, pairwise),
                             meanparam=None, covparam=gnp.asarray(th), meantype=cases.meantype_of(kind))

    def crit_for(model, kind):
        if kind == "zero":
            return lambda param, xb, zb: model.negative_log_likelihood_zero_mean(param, xb, zb)
        return lambda param, xb, zb: model.negative_log_restricted_likelihood(param, xb, zb)

    for name, n, bs, d, p, kind, seed in MINIBATCH_CASES:
        x, z, _ = cases.data(n, d, seed)
        th = cases.theta(d, seed)
        model = model_for(kind, p, th)
        loader = [(torch.as_tensor(x[i:i + bs]), torch.as_tensor(z[i:i + bs])) for i in range(0, n, bs)]
        for red in ("mean", "sum"):
            c = gnp.BatchDifferentiableSelectionCriterion(crit_for(model, kind), loader, reduction=red)
            v = c.evaluate_pre_grad(th)
            g = c.gradient(th)
            out[f"{name}/{red}/value"] = np.float64(v)
            out[f"{name}/{red}/grad"] = np.asarray(g.detach().cpu().numpy(), dtype=np.float64)
            out[f"{name}/{red}/nograd"] = np.float64(c.evaluate_no_grad(th))
        out[name + "/x"], out[name + "/z"], out[name + "/theta"] = x, z, th

    for name, n, d, p, kind, N, seed in PARTICLE_CASES:
        x, z, _ = cases.data(n, d, seed)
        th0 = cases.theta(d, seed)
        TH = th0 + np.random.default_rng(seed + 1).uniform(-1.0, 1.0, size=(N, 1 + d))
        model = model_for(kind, p, th0)
        xg, zg = gnp.asarray(x), gnp.asarray(z)
        crit = crit_for(model, kind)
        vals, grads = [], []
        for i in range(N):
            v, g = gnp.value_and_grad(lambda t: crit(t, xg, zg), gnp.asarray(TH[i]))
            vals.append(float(v))
            grads.append(np.asarray(g.detach().cpu().numpy(), dtype=np.float64))
        out[name + "/x"], out[name + "/z"], out[name + "/TH"] = x, z, TH
        out[name + "/vals"], out[name + "/grads"] = np.array(vals), np.array(grads)

    path = os.path.join(ROOT, "tests", "golden", "reference_minibatch.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, sorted(k for k in out if k.endswith("value") or k.endswith("vals")))


if __name__ == "__main__":
    main()
