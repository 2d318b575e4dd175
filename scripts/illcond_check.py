"""Agreement with the CPU oracle on ill-conditioned covariance matrices (nugget-only Matern, dense designs)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import gp_numpy as onp, gp_torch as ot

rng = np.random.default_rng(3)
for (n, d, p, rho) in [(300, 1, 3, 0.3), (500, 1, 2, 0.5), (800, 2, 2, 0.6), (1500, 3, 2, 1.0), (2000, 2, 1, 0.8)]:
    x = rng.uniform(size=(n, d))
    z = np.sin(3.0 * x.sum(axis=1)) + 0.05 * rng.standard_normal(n)
    th = np.concatenate(([0.0], np.full(d, -np.log(rho))))
    K = onp.maternp_covariance(x, x, p, th)
    cond = np.linalg.cond(K)
    P = np.ones((n, 1))
    try:
        vr, gr = ot.reml_value_and_grad(x, z, P, p, th)
    except Exception as e:  # noqa: BLE001
        print(n, d, p, "oracle failed", type(e).__name__)
        continue
    m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise), covparam=th)
    tp = torch.tensor(th, requires_grad=True)
    v = m.negative_log_restricted_likelihood(tp, x, z)
    if not torch.isfinite(v):
        print(n, d, p, f"cond={cond:.2e}", "GPU: not PD", "oracle value", vr)
        continue
    (g,) = torch.autograd.grad(v, tp)
    vn = onp.negative_log_restricted_likelihood(
        onp.OracleModel(lambda a, mp: np.ones((a.shape[0], 1)), lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, p, cp, pw),
                        None, th, "linear_predictor"), th, x, z)
    xt = rng.uniform(size=(50, d))
    mu, var = m.predict(x, z, xt)
    om = onp.OracleModel(lambda a, mp: np.ones((a.shape[0], 1)), lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, p, cp, pw),
                         None, th, "linear_predictor")
    mur, varr = onp.predict(om, x, z, xt)[:2]
    print(f"n={n} d={d} p={p} cond={cond:.2e}  value rel vs numpy {abs(v.item()-vn)/abs(vn):.2e}  vs torch {abs(v.item()-vr)/abs(vr):.2e}"
          f"  numpy-vs-torch {abs(vn-vr)/abs(vr):.2e}  grad rel {np.max(np.abs(g.numpy()-gr))/np.max(np.abs(gr)):.2e}"
          f"  mean err {np.max(np.abs(mu-mur)):.2e}  var err {np.max(np.abs(var-varr)):.2e}")
