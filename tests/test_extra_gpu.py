"""GPU parity for the rows completed in round 2, against goldens from the real reference
(oracle/make_golden_extra.py -> tests/golden/reference_extra.npz):

  * full posterior covariance, return_type=1 of gpmp/core/kriging.py:170-199 (Model.kriging_predictor[_with_zero_mean])
  * multi-start REML selection on the batched value+gradient sweep (restarts of kernel/parameter_selection.py:128-276)
  * sample_paths check_result / method handling (gpmp/core/sample_paths.py:18-63)
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, relerr_norm
from oracle import cases
from oracle.make_golden_extra import POSTCOV_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gp():
    import gpmp_b200

    assert torch.cuda.is_available(), "these tests need the B200"
    gpmp_b200._abi.lib()
    return gpmp_b200


def _extra(case):
    z = np.load(os.path.join(GOLDEN_DIR, "reference_extra.npz"))
    pre = case + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


def _cov(gp, p):
    return lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise)


@pytest.mark.parametrize("name", POSTCOV_CASES)
def test_posterior_covariance_matrix(gp, name):
    _, n, m, d, p, kind, noise, seed = next(c for c in cases.PRED_CASES if c[0] == name)
    g = _extra("postcov_" + name)
    x, z, xt = cases.data(n, d, seed, m)
    th = g["theta"]
    model = gp.core.Model(cases.mean_fn(kind, gp.num), _cov(gp, p), None, th, cases.meantype_of(kind))
    f = model.kriging_predictor_with_zero_mean if kind == "zero" else model.kriging_predictor
    lam, C = f(x, xt, 1)
    lam0, v = f(x, xt, 0)
    lam1, none = f(x, xt, -1)
    assert none is None and lam.shape == (n, m) and C.shape == (m, m)
    s2 = float(np.exp(th[0]))
    C, v = C.cpu().numpy(), v.cpu().numpy()
    ec = float(np.max(np.abs(C - g["cov"])) / s2)
    ev = float(np.max(np.abs(v - g["var"])) / s2)
    el = relerr_norm(lam.cpu().numpy(), g["lam"])
    print(f"[parity] posterior covariance {name}: cov {ec:.2e}, var {ev:.2e} (relative to sigma2), lambda {el:.2e}")
    assert ec <= 1e-9 and ev <= 1e-9 and el <= 1e-7
    assert np.max(np.abs(np.diag(C) - v)) / s2 <= 1e-12
    assert np.max(np.abs(C - C.T)) / s2 <= 1e-12
    with pytest.raises(ValueError):
        f(x, xt, 2)


def test_multistart_reml_matches_reference_restarts(gp):
    """Four restarts advanced together, one batched value+gradient sweep per line-search trial: every restart must
    reach the optimum the reference's SLSQP reaches from the same start (same basin), within the reference's own
    stopping tolerance (ftol = 1e-6)."""
    g = _extra("multistart_n64")
    name, n, d, p, kind, noise, seed = next(c for c in cases.LIK_CASES if c[0] == "lik_n64_d2_p2_const")
    x, z, _ = cases.data(n, d, seed)
    model = gp.core.Model(cases.mean_fn("const", gp.num), _cov(gp, p), None, None)
    best, fbest, info = gp.kernel.multistart_reml(model, x, z, p, g["starts"])
    err = (info["values"] - g["funs"]) / np.maximum(1.0, np.abs(g["funs"]))
    print(f"[parity] multi-start REML (4 restarts, {info['sweeps']} sweeps): final values {info['values']} vs "
          f"reference {g['funs']}; iterations {info['iterations']}")
    # never worse than the reference's end point beyond its stopping tolerance; the same optimum where both converge
    assert np.all(err <= 2e-6)
    assert abs(fbest - g["funs"].min()) <= 2e-6 * max(1.0, abs(g["funs"].min()))
    assert info["sweeps"] < 4 * 60, "restarts were not advanced together"
    v, gr = gp.num.value_and_grad(lambda t: model.negative_log_restricted_likelihood(t, x, z), best)
    assert abs(float(v) - fbest) <= 1e-9 * max(1.0, abs(fbest))


def test_sample_paths_method_and_check_result(gp):
    n, d = 90, 2
    x, _, _ = cases.data(n, d, 71)
    th = cases.theta(d, 71)
    model = gp.core.Model(None, _cov(gp, 2), None, th, "zero")
    zs = model.sample_paths(x, 7, method="chol", check_result=True)
    assert tuple(zs.shape) == (n, 7) and bool(torch.isfinite(zs).all())
    zs2 = model.sample_paths(x, 3, method="svd")
    assert tuple(zs2.shape) == (n, 3) and bool(torch.isfinite(zs2).all())
    with pytest.raises(ValueError):
        model.sample_paths(x, 3, method="qr")
    # a covariance that is not positive definite: raises when checked, returns (non-finite paths) when not
    bad = gp.core.Model(None, lambda a, b, cp, pairwise=False: -gp.kernel.maternp_covariance(a, b, 2, cp, pairwise),
                        None, th, "zero")
    with pytest.raises(torch.linalg.LinAlgError):
        bad.sample_paths(x, 2, check_result=True)
    out = bad.sample_paths(x, 2, check_result=False)
    assert tuple(out.shape) == (n, 2)


@pytest.mark.parametrize("name", ["svd_dup_n200_d2_p2", "svd_dense_n300_d1_p2", "svd_pd_n120_d3_p1"])
def test_sample_paths_svd_route_matches_reference(gp, name):
    """sample_paths(method="svd") (core/sample_paths.py:50-58): the symmetric square root U sqrt(s) V^T of K(xt, xt)
    times fixed normals, against the reference's own lines run on GPmp (oracle/make_golden_svd.py).  The two singular
    cases (duplicated points: the Cholesky route fails; a dense 1-d design, cond ~ 5e17) carry sqrt(eps)-sized noise
    in the null space in BOTH implementations, so they are held to 2e-6 of the scale of the paths and to
    C C^T = K at 2e-7 (CPU emulation of the same iteration: 1e-7 / 2e-8); the well-conditioned case to 1e-10 / 1e-12."""
    z = np.load(os.path.join(GOLDEN_DIR, "reference_svd.npz"))
    g = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}
    x, th, p, normals = g["x"], g["theta"], int(g["p"]), g["normals"]
    model = gp.core.Model(None, _cov(gp, p), None, th, "zero")
    zs = model.sample_paths_from_normals(x, normals, method="svd").cpu().numpy()
    scale = float(np.max(np.abs(g["zsim"])))
    err = float(np.max(np.abs(zs - g["zsim"]))) / scale
    singular = float(g["s"][0] / max(g["s"][-1], 1e-300)) > 1e12
    # the root itself: C = paths of the identity; C C^T must reproduce K
    C = model.sample_paths_from_normals(x, np.eye(x.shape[0]), method="svd").cpu().numpy()
    K = gp.kernel.maternp_covariance(gp.num.asarray(x), gp.num.asarray(x), p, th).cpu().numpy()
    back = float(np.max(np.abs(C @ C.T - K)) / np.max(np.abs(K)))
    print(f"[parity] {name}: paths vs reference {err:.2e} of scale, C C^T vs K {back:.2e}, "
          f"C vs reference root {np.max(np.abs(C - g['C'])) / np.max(np.abs(g['C'])):.2e}")
    assert err <= (2e-6 if singular else 1e-10)
    assert back <= (2e-7 if singular else 1e-12)
    assert np.max(np.abs(C - C.T)) <= (1e-6 if singular else 1e-12) * np.max(np.abs(C))
    if name.startswith("svd_dup"):
        with pytest.raises(torch.linalg.LinAlgError):
            model.sample_paths_from_normals(x, normals, method="chol", check_result=True)
