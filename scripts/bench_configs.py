"""Secondary measurements for the other BASELINE configs (not the judged bench line):
config 2 (n=2000, d=6, user-composed noisy kernel: per-evaluation latency), config 4 (8192 particles x n=512,
d=4: sweep time), config 5 (n=32768, d=10: REML value+grad on one GPU, chunked predict throughput)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

which = sys.argv[1:] or ["2", "4", "5"]
gnp = gp.num
out = {}


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


const_mean = lambda x_, mp: gnp.ones((x_.shape[0], 1))

if "2" in which:
    n, d, p = 2000, 6, 2
    x, z, _ = cases.data(n, d, 5)
    th = cases.theta(d, 5, noise=True)

    def noisy(a, b, cp, pairwise=False):
        s2, t2, lir = torch.exp(cp[0]), torch.exp(cp[1]), cp[2:]
        if b is a or b is None:
            if pairwise:
                return s2 * gnp.ones((a.shape[0],))
            return s2 * gp.kernel.maternp_kernel(p, gnp.scaled_distance(lir, a, a)) + t2 * gnp.eye(a.shape[0])
        if pairwise:
            return s2 * gp.kernel.maternp_kernel(p, gnp.scaled_distance_elementwise(lir, a, b))
        return s2 * gp.kernel.maternp_kernel(p, gnp.scaled_distance(lir, a, b))

    m = gp.core.Model(const_mean, noisy)
    xd, zd = gnp.asarray(x), gnp.asarray(z)

    def vg():
        tp = torch.tensor(th, requires_grad=True)
        v = m.negative_log_restricted_likelihood(tp, xd, zd)
        torch.autograd.grad(v, tp)

    out["config2_composable_value_grad_ms"] = 1e3 * timeit(vg, reps=10)
    mf = gp.core.Model(const_mean, lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
    thf = np.delete(th, 1)

    def vgf():
        tp = torch.tensor(thf, requires_grad=True)
        v = mf.negative_log_restricted_likelihood(tp, xd, zd)
        torch.autograd.grad(v, tp)

    out["config2_fused_value_grad_ms"] = 1e3 * timeit(vgf, reps=10)

if "4" in which:
    n, d, p, N = 512, 4, 2, 8192
    x, z, _ = cases.data(n, d, 77)
    th0 = cases.theta(d, 77)
    TH = th0 + np.random.default_rng(1).uniform(-2.0, 2.0, size=(N, d + 1))
    m = gp.core.Model(const_mean, lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
    crit = gp.batched.BatchedCriterion(m, x, z, p, kind="reml")
    thd = gnp.asarray(TH)
    t_dev = timeit(lambda: crit.values_device(thd), reps=5)
    t_e2e = timeit(lambda: crit(TH), reps=5)
    vals = crit(TH)
    out["config4_sweep_ms_device"] = 1e3 * t_dev
    out["config4_sweep_ms_e2e"] = 1e3 * t_e2e
    out["config4_particle_evals_per_s"] = N / t_e2e
    out["config4_tflops"] = N * n**3 / 3 / t_dev / 1e12
    out["config4_finite_fraction"] = float(np.isfinite(vals).mean())

if "5" in which:
    n, d, p = 32768, 10, 2
    x, z, _ = cases.data(n, d, 9)
    th = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
    m = gp.core.Model(const_mean, lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise), covparam=th)
    xd, zd = gnp.asarray(x), gnp.asarray(z)

    def val():
        with torch.no_grad():
            return m.negative_log_restricted_likelihood(th, xd, zd)

    def vg():
        tp = torch.tensor(th, requires_grad=True)
        v = m.negative_log_restricted_likelihood(tp, xd, zd)
        (g,) = torch.autograd.grad(v, tp)
        return v, g

    t_v = timeit(val, reps=2, warm=1)
    t_vg = timeit(vg, reps=2, warm=1)
    v, g = vg()
    out["config5_value_s"] = t_v
    out["config5_value_tflops"] = n**3 / 3 / t_v / 1e12
    out["config5_value_grad_s"] = t_vg
    out["config5_value_grad_tflops"] = float(n) ** 3 / t_vg / 1e12
    out["config5_value"] = v.item()
    # closed form for d/d log sigma2 as a full-size self-check
    quad = m.norm_k_sqrd(xd, zd, th).item()
    out["config5_grad0_vs_closed_form"] = abs(g[0].item() - 0.5 * ((n - 1) - quad))
    mt = 65536
    xt = np.random.default_rng(10).uniform(size=(mt, d))
    xtd = gnp.asarray(xt)
    t_p = timeit(lambda: m.predict(xd, zd, xtd, convert_out=False), reps=1, warm=1)
    out["config5_predict_points_per_s"] = mt / t_p
    out["config5_predict_tflops"] = (n**3 / 3 + float(n) * n * mt) / t_p / 1e12

print(json.dumps(out, indent=1))
