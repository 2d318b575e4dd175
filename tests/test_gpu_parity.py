"""GPU parity tests: the CUDA path (through the C-ABI, via gpmp_b200) against the committed golden
vectors of the real reference (tests/golden/) and against the CPU oracle on seeded inputs.

Tolerances are BASELINE.json's: covariance rel-err <= 1e-12, logL and gradient <= 1e-8 relative,
predictive mean / variance <= 1e-9 (relative to the prior scale sigma).
"""
import numpy as np
import pytest
import torch

from conftest import relerr, relerr_norm
from oracle import cases, gp_numpy as onp

pytestmark = pytest.mark.gpu

TOL_COV = 1e-12
TOL_LIK = 1e-8
TOL_PRED = 1e-9


@pytest.fixture(scope="module")
def gp():
    import gpmp_b200

    assert torch.cuda.is_available(), "these tests need the B200"
    gpmp_b200._abi.lib()  # fail loudly if the native library is missing
    return gpmp_b200


def _cov_callable(gp, p, noise):
    if not noise:
        return lambda x, y, cp, pairwise=False: gp.kernel.maternp_covariance(x, y, p, cp, pairwise)
    gnp = gp.num

    def k(x, y, cp, pairwise=False):
        # examples/gpmp_example07_nd_regression.py:95-130: sigma2 k_p(D) + tau2 I composed by the user
        s2, t2, lir = torch.exp(cp[0]), torch.exp(cp[1]), cp[2:]
        if y is x or y is None:
            if pairwise:
                return s2 * gnp.ones((x.shape[0],))
            D = gnp.scaled_distance(lir, x, x)
            return s2 * gp.kernel.maternp_kernel(p, D) + t2 * gnp.eye(x.shape[0])
        if pairwise:
            return s2 * gp.kernel.maternp_kernel(p, gnp.scaled_distance_elementwise(lir, x, y))
        return s2 * gp.kernel.maternp_kernel(p, gnp.scaled_distance(lir, x, y))

    return k


def _model(gp, kind, p, noise, th):
    mp = cases.MEANPARAM if kind == "param" else None
    return gp.core.Model(cases.mean_fn(kind, gp.num), _cov_callable(gp, p, noise), mp, th, cases.meantype_of(kind))


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c[0] for c in cases.COV_CASES])
def test_covariance(gp, golden_np, case):
    g = golden_np(case)
    x, y, th, p = g["x"], g["y"], g["theta"], int(g["p"])
    gnp = gp.num
    xd, yd = gnp.asarray(x), gnp.asarray(y)
    k = min(len(x), len(y))
    D = gnp.scaled_distance(th[1:], xd, yd).cpu().numpy()
    assert relerr_norm(D, g["D"]) <= TOL_COV
    Dxx = gnp.scaled_distance(th[1:], xd, xd).cpu().numpy()
    assert np.all(np.diag(Dxx) == 0.0) and relerr_norm(Dxx, g["Dxx"]) <= TOL_COV
    De = gnp.scaled_distance_elementwise(th[1:], gnp.asarray(x[:k]), gnp.asarray(y[:k])).cpu().numpy()
    assert relerr_norm(De, g["De"]) <= TOL_COV
    Kii = gp.kernel.maternp_covariance(xd, xd, p, th).cpu().numpy()
    assert relerr(Kii, g["Kii"]) <= TOL_COV
    assert np.array_equal(Kii, Kii.T)
    Kn = gp.kernel.maternp_covariance(xd, None, p, th).cpu().numpy()
    assert relerr(Kn, g["Kii_none"]) <= TOL_COV
    Kit = gp.kernel.maternp_covariance(xd, yd, p, th).cpu().numpy()
    assert relerr(Kit, g["Kit"]) <= TOL_COV
    assert relerr(gp.kernel.maternp_covariance(xd, None, p, th, True).cpu().numpy(), g["Kii_pw"]) <= TOL_COV
    pw = gp.kernel.maternp_covariance(gnp.asarray(x[:k]), gnp.asarray(y[:k]), p, th, True).cpu().numpy()
    assert relerr(pw, g["Kit_pw"]) <= TOL_COV
    kh = gp.kernel.maternp_kernel(p, gnp.asarray(g["h"])).cpu().numpy()
    ok = ~np.isnan(g["kh"])
    assert np.array_equal(np.isnan(kh), np.isnan(g["kh"]))
    assert relerr_norm(kh[ok], g["kh"][ok]) <= TOL_COV


def test_covariance_empty_and_ragged(gp):
    gnp = gp.num
    x = gnp.asarray(np.random.default_rng(0).uniform(size=(65, 3)))
    th = np.array([0.1, 0.2, -0.3, 0.4])
    assert gp.kernel.maternp_covariance(x[:0], x[:0].clone(), 2, th).shape == (0, 0)
    K = gp.kernel.maternp_covariance(x, x[:1].clone(), 2, th)
    assert K.shape == (65, 1)
    ref = onp.maternp_covariance(x.cpu().numpy(), x[:1].cpu().numpy(), 2, th)
    assert relerr(K.cpu().numpy(), ref) <= TOL_COV


def test_covariance_param_gradient(gp):
    """vjp of the covariance op (dK regenerated per tile) against torch-CPU autograd of the oracle."""
    from oracle import gp_torch as ot

    rng = np.random.default_rng(5)
    x = rng.uniform(size=(70, 3))
    th = np.array([0.3, 0.1, -0.2, 0.5])
    G = rng.standard_normal((70, 70))
    for p in (0, 1, 2, 4):
        tp = torch.tensor(th, requires_grad=True)
        K = gp.kernel.maternp_covariance(gp.num.asarray(x), None, p, tp)
        (K * gp.num.asarray(G)).sum().backward()
        tr = torch.tensor(th, requires_grad=True)
        Kr = ot.matern_cov_ii(torch.tensor(x), p, tr)
        (Kr * torch.tensor(G)).sum().backward()
        assert relerr_norm(tp.grad.numpy(), tr.grad.numpy()) <= 1e-10, p


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(128, 128, 128), (300, 200, 70), (257, 129, 515), (1, 5, 3), (640, 384, 1024)])
def test_gemm_nt(gp, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A, B, C0 = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    ops = gp.ops
    Ad, Bd, Cd = ops.padded(ops.to_device(A)), ops.padded(ops.to_device(B)), ops.padded(ops.to_device(C0))
    ops.gemm_nt(Ad, Bd, C_out=Cd, alpha=-0.5, beta=2.0)
    ref = -0.5 * A @ B.T + 2.0 * C0
    assert relerr_norm(Cd.cpu().numpy(), ref) <= 1e-13


@pytest.mark.parametrize("M,N,K,lower,tri", [(2085, 1900, 515, False, 0), (2432, 2432, 130, True, 0), (2432, 2432, 704, True, 0),
                                             (2048, 2048, 2048, False, 1), (2300, 2300, 2300, True, 1),
                                             (2048, 1700, 16, False, 0)])
def test_gemm_nt_machine_filling_shapes(gp, M, N, K, lower, tri):
    """Shapes with at least one 128 x 128 tile per SM take the TMA / mbarrier kernel (gemm_tma.cu): ragged edges are
    zero-filled by the TMA unit, lower tile lists and the triangular K trimming follow the 128-tile grid, alpha /
    beta are applied in the epilogue.  tri = 1 is KR_FROM_ROW: the k loop starts at the tile's first row (operand A
    upper triangular on the 128-tile grid)."""
    rng = np.random.default_rng(M + N + K)
    A, B, C0 = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    ops = gp.ops
    if tri == 1:  # only k >= 128 * floor(row / 128) contributes: make the reference see the same thing
        rows = (np.arange(M) // 128 * 128)[:, None]
        A = np.where(np.arange(K)[None, :] >= rows, A, 0.0)
        Aj = A + np.where(np.arange(K)[None, :] < rows, 7.0, 0.0)  # junk left of the tile grid must be skipped
    else:
        Aj = A
    Ad, Bd, Cd = ops.padded(ops.to_device(Aj)), ops.padded(ops.to_device(B)), ops.padded(ops.to_device(C0))
    l0 = gp._abi.launch_count()
    ops.gemm_nt(Ad, Bd, C_out=Cd, alpha=-0.5, beta=2.0, tri=tri, lower=lower)
    assert gp._abi.launch_count() == l0 + 1
    out = Cd.cpu().numpy()
    ref = -0.5 * A @ B.T + 2.0 * C0
    if lower:
        # tiles above the diagonal of the 128-tile grid are never visited; everything on or below the diagonal of the
        # 64-tile grid always is (the 64 x 64 kernel serves short k loops, the 128 x 128 TMA kernel long ones)
        r, c = np.arange(M)[:, None], np.arange(N)[None, :]
        untouched, visited = c // 128 > r // 128, c // 64 <= r // 64
        assert np.array_equal(out[untouched], C0[untouched])
        out, ref = np.where(visited, out, 0.0), np.where(visited, ref, 0.0)
    assert relerr_norm(out, ref) <= 1e-13


@pytest.mark.parametrize("sm_first", [0, 40, 147, 1000])
@pytest.mark.parametrize("M,N,K,lower,tri,pairs", [(1100, 900, 300, False, 0, 1), (1024, 1024, 1024, True, 1, 1),
                                                   (512, 512, 512, False, 1, 3), (2048, 2048, 130, True, 0, 1)])
def test_gemm_nt_persistent_any_placement(gp, M, N, K, lower, tri, pairs, sm_first):
    """The persistent form of the NT GEMM (gemm.cu gemm_nt_persist_kernel: the levels of T = L^-1 that run under the
    tail of a factorisation): CTAs on SMs below sm_first leave at once, the others draw (tile, pair) indices from a
    counter; when half the grid has arrived and nobody has drawn a tile yet, the CTAs arriving from then on work
    wherever they land, and the last CTA to leave finishes what is left.  The result must not depend on placement:
    every CTA admitted (0), most (40), a single SM (147), none at all (1000: the stranded-grid rule does the work)."""
    import ctypes as C

    rng = np.random.default_rng(M + N + K + pairs)
    ops = gp.ops
    A = rng.standard_normal((pairs, M, K))
    B = rng.standard_normal((pairs, N, K))
    C0 = rng.standard_normal((pairs, M, N))
    if tri == 1:
        rows = (np.arange(M) // 128 * 128)[:, None]
        keep = np.arange(K)[None, :] >= rows
        Aj = np.where(keep, A, 7.0)  # junk left of the tile grid must be skipped
        A = np.where(keep, A, 0.0)
    else:
        Aj = A
    ldk, ldn = (K + 15) // 16 * 16, (N + 15) // 16 * 16
    Ad = torch.zeros((pairs, M, ldk), dtype=torch.float64, device="cuda")
    Bd = torch.zeros((pairs, N, ldk), dtype=torch.float64, device="cuda")
    Cd = torch.zeros((pairs, M, ldn), dtype=torch.float64, device="cuda")
    Ad[:, :, :K] = torch.as_tensor(Aj)
    Bd[:, :, :K] = torch.as_tensor(B)
    Cd[:, :, :N] = torch.as_tensor(C0)
    lib = gp._abi.lib()
    fn = lib.gpmp_debug_gemm_nt_persist
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int,
                   C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_longlong,
                   C.c_longlong, C.c_int, C.c_void_p]
    for _ in range(2):  # twice: the counter pair must come back reset
        Cd[:, :, :N] = torch.as_tensor(C0)
        rc = fn(Ad.data_ptr(), ldk, Bd.data_ptr(), ldk, Cd.data_ptr(), ldn, M, N, K, -0.5, 2.0, tri, int(lower), pairs,
                M * ldk, N * ldk, M * ldn, sm_first, gp._abi.stream_ptr())
        assert rc == 0
        out = Cd[:, :, :N].cpu().numpy()
        ref = -0.5 * np.einsum("pmk,pnk->pmn", A, B) + 2.0 * C0
        if lower:
            r, c = np.arange(M)[:, None], np.arange(N)[None, :]
            untouched, visited = (c // 64 > r // 64)[None], (c // 64 <= r // 64)[None]
            assert np.array_equal(out[np.broadcast_to(untouched, out.shape)], C0[np.broadcast_to(untouched, out.shape)])
            out, ref = np.where(visited, out, 0.0), np.where(visited, ref, 0.0)
        assert relerr_norm(out, ref) <= 1e-13


@pytest.mark.parametrize("n", [1, 6, 64, 127, 128, 129, 300, 700, 1500, 2500])
def test_potrf_potri_trsm(gp, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(size=(n, 3))
    K = onp.maternp_covariance(x, x, 2, np.array([0.2, 1.0, 1.2, 0.8])) + 1e-6 * np.eye(n)
    B = rng.standard_normal((5, n))
    ops = gp.ops
    fac = ops.potrf(ops.to_device(K), extra_rows=ops.to_device(B))
    L = np.linalg.cholesky(K)
    Lg = fac.lower().cpu().numpy()
    assert relerr_norm(Lg, L) <= 1e-10
    # extra rows ride along: B L^-T
    from scipy.linalg import solve_triangular

    W = solve_triangular(L, B.T, lower=True).T
    assert relerr_norm(fac.A[n:, :n].cpu().numpy(), W) <= 1e-9
    # mirrored upper tiles hold L^T
    full = fac.A[:n, :n].cpu().numpy()
    t = 128
    for bi in range(0, n, t):
        for bj in range(bi + t, n, t):
            assert np.array_equal(full[bi:bi + t, bj:bj + t], Lg[bj:bj + t, bi:bi + t].T)
    Kinv, Tlo, Tup = ops.potri(fac)
    Ki = np.linalg.inv(K)
    assert relerr_norm(np.tril(Kinv[:, :n].cpu().numpy()), np.tril(Ki)) <= 1e-7
    Tl = np.tril(Tlo[:, :n].cpu().numpy())
    assert relerr_norm(Tl, np.linalg.inv(L)) <= 1e-8
    assert relerr_norm(np.triu(Tup[:, :n].cpu().numpy()), Tl.T) <= 1e-15
    R = ops.padded(ops.to_device(B))
    ops.trsm_rows(fac, R, trans=0)
    assert relerr_norm(R.cpu().numpy(), W) <= 1e-9
    ops.trsm_rows(fac, R, trans=1)
    assert relerr_norm(R.cpu().numpy(), np.linalg.solve(K, B.T).T) <= 1e-7


def test_potrf_not_positive_definite(gp):
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(torch.linalg.LinAlgError):
        gp.num.cholesky(A)
    fac = gp.ops.potrf(gp.ops.to_device(A), check_pd=False)
    assert int(fac.info.item()) == 151


def test_gnp_cholesky_family(gp):
    rng = np.random.default_rng(3)
    n = 333
    x = rng.uniform(size=(n, 2))
    K = onp.maternp_covariance(x, x, 1, np.array([0.0, 1.5, 1.5])) + 1e-8 * np.eye(n)
    b = rng.standard_normal(n)
    B = rng.standard_normal((n, 4))
    gnp = gp.num
    xs, L = gnp.cholesky_solve(K, b)
    assert xs.shape == (n, 1)
    assert relerr_norm(xs.cpu().numpy()[:, 0], np.linalg.solve(K, b)) <= 1e-8
    assert relerr_norm(L.cpu().numpy(), np.linalg.cholesky(K)) <= 1e-10
    X2, _ = gnp.cholesky_solve(K, B)
    assert relerr_norm(X2.cpu().numpy(), np.linalg.solve(K, B)) <= 1e-8
    assert relerr_norm(gnp.cholesky_inv(K).cpu().numpy(), np.linalg.inv(K)) <= 1e-7
    y = gnp.solve_triangular(L, B, lower=True)
    assert relerr_norm(y.cpu().numpy(), np.linalg.solve(np.linalg.cholesky(K), B)) <= 1e-9
    c, low = gnp.cho_factor(K, lower=True)
    assert relerr_norm(gnp.cho_solve((c, low), b).cpu().numpy(), np.linalg.solve(K, b)) <= 1e-8


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c[0] for c in cases.LIK_CASES])
def test_likelihoods_and_gradients(gp, golden_np, golden_t, case):
    name, n, d, p, kind, noise, seed = next(c for c in cases.LIK_CASES if c[0] == case)
    gn, gt = golden_np(case), golden_t(case)
    x, z, th = gn["x"], gn["z"], gn["theta"]
    m = _model(gp, kind, p, noise, th)
    # values against the NumPy-backend reference
    v = m.negative_log_likelihood_zero_mean(th, x, z)
    assert relerr(v.item(), gn["nll_zero"]) <= TOL_LIK
    assert relerr(m.norm_k_sqrd_with_zero_mean(x, z, th).item(), gn["norm_zero"]) <= TOL_LIK
    zkz, ki1, kiz = m.k_inverses(x, z, th)
    assert relerr(zkz.item(), gn["kinv_zkz"]) <= TOL_LIK
    assert relerr_norm(ki1.cpu().numpy(), gn["kinv_1"]) <= 1e-7 and relerr_norm(kiz.cpu().numpy(), gn["kinv_z"]) <= 1e-7
    # gradients against the torch-backend reference (autograd through every op)
    tp = torch.tensor(th, requires_grad=True)
    v = m.negative_log_likelihood_zero_mean(tp, x, z)
    (g,) = torch.autograd.grad(v, tp)
    assert g.device.type == "cpu" and g.shape == tp.shape
    assert relerr_norm(g.numpy(), gt["nll_zero_grad"]) <= TOL_LIK
    if "reml" in gn:
        tp = torch.tensor(th, requires_grad=True)
        v = m.negative_log_restricted_likelihood(tp, x, z)
        assert v.ndim == 0 and v.grad_fn is not None
        assert relerr(v.item(), gn["reml"]) <= TOL_LIK
        (g,) = torch.autograd.grad(v, tp)
        assert relerr_norm(g.numpy(), gt["reml_grad"]) <= TOL_LIK
        assert relerr(m.norm_k_sqrd(x, z, th).item(), gn["norm_k"]) <= TOL_LIK
    if "nll_param" in gn:
        full = torch.tensor(np.concatenate((cases.MEANPARAM, th)), requires_grad=True)
        v = m.negative_log_likelihood(full[:2], full[2:], x, z)
        assert relerr(v.item(), gn["nll_param"]) <= TOL_LIK
        (g,) = torch.autograd.grad(v, full)
        assert relerr_norm(g.numpy(), gt["nll_param_grad"]) <= TOL_LIK


@pytest.mark.parametrize("case", [c[0] for c in cases.LIK_CASES])
def test_loo(gp, golden_np, case):
    name, n, d, p, kind, noise, seed = next(c for c in cases.LIK_CASES if c[0] == case)
    g = golden_np(case)
    m = _model(gp, kind, p, noise, g["theta"])
    zloo, s2, e = m.loo(g["x"], g["z"], convert_out=True)
    assert relerr_norm(zloo, g["loo_z"]) <= 1e-7
    assert relerr_norm(s2, g["loo_s2"]) <= 1e-7
    assert relerr_norm(e, g["loo_e"]) <= 1e-7


def test_likelihood_not_pd_gives_inf(gp):
    x = np.random.default_rng(0).uniform(size=(40, 2))
    z = np.arange(40.0)
    m = gp.core.Model(cases.mean_fn("const", gp.num),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise))
    th = torch.tensor([-800.0, 0.0, 0.0], requires_grad=True)  # sigma2 = exp(-800) = 0: K == 0, not PD
    v = m.negative_log_restricted_likelihood(th, x, z)
    assert torch.isinf(v) and v > 0
    val, grad = gp.num.value_and_grad(lambda t: m.negative_log_restricted_likelihood(t, x, z), th)
    assert torch.isinf(val) and torch.all(grad == 0)
    # same through the three-stream look-ahead path (n large enough for several column groups)
    xb = np.random.default_rng(1).uniform(size=(3000, 2))
    zb = np.sin(xb.sum(1))
    big = m.negative_log_restricted_likelihood(torch.tensor([-800.0, 0.0, 0.0]), xb, zb)
    assert torch.isinf(big) and big > 0
    # and a failure in the middle of a user-composed matrix: one negative diagonal entry at row 1700
    Kbad = torch.eye(3000, dtype=torch.float64, device="cuda")
    Kbad[1700, 1700] = -1.0
    mb = gp.core.Model(cases.mean_fn("const", gp.num), lambda a, b, cp, pairwise=False: Kbad * torch.exp(cp[0]).item())
    assert torch.isinf(mb.negative_log_restricted_likelihood(torch.tensor([0.0]), xb, zb))


def test_reml_selection_example02_end_to_end(gp, golden_t):
    """BASELINE config 1: the SciPy SLSQP loop of select_parameters_with_reml
    (kernel/parameter_selection.py:128-276: same start, +-10 box, ftol=1e-6, eps=1e-8, best-seen tracking)
    driven by this library's criterion and gradient must land on the reference's selected parameters, and
    predict from there must reproduce its posterior mean / variance."""
    from scipy.optimize import minimize

    g = golden_t("select_reml_example02")
    x, z, xt, p, th0 = g["x"], g["z"], g["xt"], int(g["p"]), g["covparam0"]
    m = gp.core.Model(cases.mean_fn("const", gp.num),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
    crit = gp.num.DifferentiableSelectionCriterion(
        lambda p_, x_, z_: m.negative_log_restricted_likelihood(p_, x_, z_), x, z)
    best = {"J": np.inf, "p": None}

    def fun(pv):
        J = crit.evaluate_pre_grad(pv)
        if J < best["J"]:
            best["J"], best["p"] = J, pv.copy()
        return J

    r = minimize(fun, th0, method="SLSQP", jac=lambda pv: np.asarray(crit.gradient(pv)),
                 bounds=[(v - 10.0, v + 10.0) for v in th0], options=dict(ftol=1e-6, eps=1e-8, maxiter=15000))
    assert r.success
    assert abs(best["J"] - float(g["fun"])) <= 1e-6 * max(1.0, abs(float(g["fun"])))
    assert np.max(np.abs(best["p"] - g["covparam"])) <= 1e-3
    # the initial guess (kernel/init.py:54-66, one hot-path call) reproduces the reference's starting point; GPmp's
    # own driver on top of this library is exercised in tests/test_dropin_gpu.py
    th_init = gp.kernel.anisotropic_parameters_initial_guess(m, gp.num.asarray(x), gp.num.asarray(z))
    assert np.max(np.abs(th_init - th0)) <= 1e-8 * max(1.0, np.max(np.abs(th0)))
    m.covparam = g["covparam"]  # predict at the reference's parameters: isolates the predictor from the optimiser
    mean, var = m.predict(x, z, xt)
    s2 = float(np.exp(g["covparam"][0]))
    assert np.max(np.abs(mean - g["mean"])) / max(np.sqrt(s2), np.max(np.abs(g["mean"]))) <= 1e-8
    assert np.max(np.abs(var - g["var"])) / s2 <= 1e-8


def test_selection_criterion_lifecycle(gp, golden_np, golden_t):
    """evaluate_pre_grad / gradient / evaluate_no_grad as SciPy drives them (torch_backend.py:547-604)."""
    case = "lik_n500_d8_p2_const"
    gn, gt = golden_np(case), golden_t(case)
    m = _model(gp, "const", 2, False, gn["theta"])
    crit = gp.num.DifferentiableSelectionCriterion(
        lambda p_, x_, z_: m.negative_log_restricted_likelihood(p_, x_, z_), gn["x"], gn["z"])
    f = crit.evaluate_pre_grad(gn["theta"])
    assert isinstance(f, float) and relerr(f, gn["reml"]) <= TOL_LIK
    g = crit.gradient(gn["theta"])
    assert relerr_norm(np.asarray(g), gt["reml_grad"]) <= TOL_LIK
    f2 = crit.evaluate_no_grad(gn["theta"])
    assert relerr(float(f2), gn["reml"]) <= TOL_LIK


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c[0] for c in cases.PRED_CASES])
def test_predict_and_conditioning(gp, golden_np, case):
    name, n, mt, d, p, kind, noise, seed = next(c for c in cases.PRED_CASES if c[0] == case)
    g = golden_np(case)
    x, z, xt, th = g["x"], g["z"], g["xt"], g["theta"]
    m = _model(gp, kind, p, noise, th)
    sigma2 = float(np.exp(th[0]))
    mean, var, lam = m.predict(x, z, xt, return_lambdas=True)
    assert isinstance(mean, np.ndarray) and mean.shape == (mt,) and lam.shape == (n, mt)
    scale = max(float(np.max(np.abs(g["mean"]))), np.sqrt(sigma2))
    assert np.max(np.abs(mean - g["mean"])) / scale <= TOL_PRED
    assert np.max(np.abs(var - g["var"])) / sigma2 <= TOL_PRED
    assert relerr_norm(lam.cpu().numpy(), g["lam"]) <= 1e-7
    mean2, var2 = m.predict(x, z, xt)
    assert np.array_equal(mean2, mean) or np.max(np.abs(mean2 - mean)) / scale <= 1e-12
    # conditioning by kriging on the reference's unconditional paths
    xi_ind, xt_ind = np.arange(n), n + np.arange(mt)
    if kind == "param":
        cond = m.conditional_sample_paths_parameterized_mean(g["ztsim"], x, xi_ind, z, xt, xt_ind, lam)
    else:
        cond = m.conditional_sample_paths(g["ztsim"], xi_ind, z, xt_ind, lam)
    assert cond.shape == g["cond"].shape
    assert np.max(np.abs(cond - g["cond"])) / max(1.0, float(np.max(np.abs(g["cond"])))) <= 1e-8
    # the chunked form never builds lambda_t: W (L^-1 delta)^T per chunk of xt
    fit = m.fit(x, z)
    cond2 = fit.conditional_sample_paths_chunked(g["ztsim"], xi_ind, xt, xt_ind)
    assert np.max(np.abs(cond2 - g["cond"])) / max(1.0, float(np.max(np.abs(g["cond"])))) <= 1e-8
    mean3, var3 = fit.predict(xt)
    assert np.array_equal(mean3, mean2) and np.array_equal(var3, var2)
    # the deterministic half of sample_paths: C @ normals with K(xt, xt) = C C^T
    normals = np.random.default_rng(1).standard_normal((mt, 3))
    zs = m.sample_paths_from_normals(xt, normals).cpu().numpy()
    Ktt = g["Ktt"]
    try:
        C = np.linalg.cholesky(Ktt)
        assert relerr_norm(zs, C @ normals) <= 1e-6
    except np.linalg.LinAlgError:
        pass


# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", [c[0] for c in cases.BATCH_CASES])
def test_batched_criterion(gp, golden_np, case):
    name, n, d, p, N, seed = next(c for c in cases.BATCH_CASES if c[0] == case)
    g = golden_np(case)
    m = _model(gp, "const", p, False, g["TH"][0])
    crit = gp.batched.BatchedCriterion(m, g["x"], g["z"], p, kind="reml")
    vals = crit(g["TH"])
    assert vals.shape == (N,)
    # particles are drawn in a wide box: some covariance matrices are badly conditioned, where the reference's
    # own two backends already differ by more than 1e-8; the tolerance follows cond(K)
    for i in range(N):
        cond = np.linalg.cond(onp.maternp_covariance(g["x"], g["x"], p, g["TH"][i]))
        tol = TOL_LIK if cond <= 1e9 else 1e-6
        assert relerr(vals[i], g["vals"][i]) <= tol, (i, cond)
    # small workspace -> several chunks, same answer
    crit2 = gp.batched.BatchedCriterion(m, g["x"], g["z"], p, kind="reml", max_bytes=3 << 20)
    assert np.array_equal(crit2(g["TH"]), vals)
    lp = crit.logpdf_temp(g["TH"], 2.0, lower=g["TH"].min(0) + 1e-9, upper=g["TH"].max(0) + 1.0)
    ref = onp.logpdf_temp(lambda t: vals[np.argmin(np.abs(g["TH"] - t).sum(1))], g["TH"], 2.0,
                          g["TH"].min(0) + 1e-9, g["TH"].max(0) + 1.0)
    assert np.array_equal(np.isinf(lp), np.isinf(ref))
    assert relerr(lp[~np.isinf(lp)], ref[~np.isinf(ref)]) <= 1e-12


def test_batched_matches_scalar_path_n512(gp):
    """config-4 shape (n=512, d=4): batched values == the scalar fused path, particle by particle."""
    x, z, _ = cases.data(512, 4, 77)
    rng = np.random.default_rng(78)
    th0 = cases.theta(4, 77)
    TH = th0 + rng.uniform(-1.0, 1.0, size=(24, 5))
    m = _model(gp, "const", 2, False, th0)
    vals = gp.batched.BatchedCriterion(m, x, z, 2)(TH)
    for i in range(0, 24, 5):
        v = m.negative_log_restricted_likelihood(TH[i], x, z).item()
        assert relerr(vals[i], v) <= 1e-12
    ref = onp.negative_log_restricted_likelihood(
        onp.OracleModel(cases.mean_fn("const", np), lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, 2, cp, pw),
                        None, th0, "linear_predictor"), TH[3], x, z)
    assert relerr(vals[3], ref) <= TOL_LIK


@pytest.mark.parametrize("kind,mean,d", [("ml", "zero", 3), ("reml", "linear", 4), ("reml", "linear", 8)])
def test_batched_extra_rows_variants(gp, kind, mean, d):
    """The whitening rows ride through the batched panel solves as an extra strip when there are at most 8 of
    them (zero mean: 1 row, linear d=4: 6 rows) and as ordinary panel rows otherwise (linear d=8: 10 rows):
    both must reproduce the scalar path."""
    n = 300
    x, z, _ = cases.data(n, d, 31)
    th0 = cases.theta(d, 31)
    TH = th0 + np.random.default_rng(32).uniform(-0.7, 0.7, size=(7, d + 1))
    m = _model(gp, mean, 2, False, th0)
    vals = gp.batched.BatchedCriterion(m, x, z, 2, kind=kind)(TH)
    for i in range(7):
        if kind == "ml":
            v = m.negative_log_likelihood_zero_mean(TH[i], x, z).item()
        else:
            v = m.negative_log_restricted_likelihood(TH[i], x, z).item()
        assert relerr(vals[i], v) <= 1e-10, (i, vals[i], v)


@pytest.mark.parametrize("n,mean,d,N", [(1100, "const", 3, 1), (1100, "zero", 3, 3), (2048, "const", 4, 1),
                                          (2048, "linear", 8, 3), (1500, "const", 6, 4)])
def test_batched_large_n_matches_scalar(gp, n, mean, d, N):
    """Batched sweeps beyond one column group (n > 1024: NB > 128, in-group K=128 updates combined with the
    extra-row strip, block inverses with batch > 1), with ONE particle (the MH loop: must not take the single-matrix
    look-ahead path, whose block-diagonal inverses need ceil(n/128) tiles) and with workspaces that force
    cap = 1 and cap = N - 1 particles in flight: every value equals the scalar fused path."""
    x, z, _ = cases.data(n, d, 300 + n)
    th0 = cases.theta(d, 300 + n)
    TH = th0 + np.random.default_rng(n + N).uniform(-0.5, 0.5, size=(N, d + 1))
    m = _model(gp, mean, 2, False, th0)
    kind = "ml" if mean == "zero" else "reml"
    f = m.negative_log_likelihood_zero_mean if mean == "zero" else m.negative_log_restricted_likelihood
    ref = np.array([f(TH[i], x, z).item() for i in range(N)])
    q = {"zero": 0, "const": 1, "linear": d + 1}[mean]
    per1 = gp._abi.lib().gpmp_criterion_batched_bytes(n, q, 1)
    per = gp._abi.lib().gpmp_criterion_batched_bytes(n, q, 2) - per1
    budgets = [None, per1] + ([per1 + (N - 2) * per] if N > 2 else [])
    for mb in budgets:
        vals = gp.batched.BatchedCriterion(m, x, z, 2, kind=kind, max_bytes=mb)(TH)
        assert relerr(vals, ref) <= 1e-10, (mb, vals, ref)
    if N > 1 and mean != "zero":
        v2, g2 = gp.batched.BatchedCriterion(m, x, z, 2, kind=kind).value_and_grad(TH[:2])
        for i in range(2):
            v, gi = gp.num.value_and_grad(lambda t: f(t, x, z), TH[i])
            assert relerr(v2[i], float(v)) <= 1e-9 and relerr_norm(g2[i], gi.cpu().numpy()) <= 1e-8


def test_batched_operand_validation(gp):
    """theta width / rank / device are checked before anything reaches the C side; an isotropic parameter row
    is expanded (and its gradient folded back) like the scalar path does."""
    n, d = 96, 3
    x, z, _ = cases.data(n, d, 5)
    th0 = cases.theta(d, 5)
    m = _model(gp, "const", 2, False, th0)
    crit = gp.batched.BatchedCriterion(m, x, z, 2)
    with pytest.raises(gp._abi.GpmpError):
        crit(np.zeros((4, d + 3)))
    iso = np.array([[0.1, 0.3], [-0.2, 0.5]])
    full = np.concatenate((iso[:, :1], np.repeat(iso[:, 1:], d, axis=1)), axis=1)
    assert np.array_equal(crit(iso), crit(full))
    vi, gi = crit.value_and_grad(iso)
    vf, gf = crit.value_and_grad(full)
    assert gi.shape == (2, 2) and relerr(vi, vf) == 0.0
    assert relerr_norm(gi[:, 1], gf[:, 1:].sum(axis=1)) <= 1e-13
    # CPU / non-contiguous tensors are moved and packed, not handed over as they are
    tht = torch.as_tensor(np.asfortranarray(full))
    assert np.array_equal(crit(tht), crit(full))


def _mb_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_minibatch.npz"))


def _mb_cases():
    from oracle.make_golden_minibatch import MINIBATCH_CASES, PARTICLE_CASES
    return MINIBATCH_CASES, PARTICLE_CASES


@pytest.mark.parametrize("case", [c[0] for c in _mb_cases()[1]])
def test_batched_value_and_grad(gp, case):
    """Per-particle value + gradient in one batched call (the SVGD loop of mcmc/svgd.py:310-313) against the
    reference's gnp.value_and_grad per particle, and against this library's scalar path."""
    name, n, d, p, kind, N, seed = next(c for c in _mb_cases()[1] if c[0] == case)
    g = _mb_golden()
    x, z, TH = g[name + "/x"], g[name + "/z"], g[name + "/TH"]
    m = _model(gp, kind, p, False, TH[0])
    crit = gp.batched.BatchedCriterion(m, x, z, p, kind="ml" if kind == "zero" else "reml")
    vals, grads = crit.value_and_grad(TH)
    assert vals.shape == (N,) and grads.shape == (N, 1 + d)
    assert relerr(vals, g[name + "/vals"]) <= TOL_LIK
    assert relerr_norm(grads, g[name + "/grads"]) <= TOL_LIK
    # values agree with the value-only sweep; gradients with the scalar autograd path (different operation
    # orders: the smooth p = 3 case has cond(K) ~ 1e10, so rounding-level differences reach 1e-9 there)
    assert relerr(vals, crit(TH)) <= TOL_LIK
    f = m.negative_log_likelihood_zero_mean if kind == "zero" else m.negative_log_restricted_likelihood
    for i in (0, N - 1):
        v, gi = gp.num.value_and_grad(lambda t: f(t, x, z), TH[i])
        assert relerr(vals[i], float(v)) <= TOL_LIK and relerr_norm(grads[i], gi.cpu().numpy()) <= TOL_LIK
    # a single row, and a small workspace (several chunks)
    v1, g1 = crit.value_and_grad(TH[2])
    assert relerr(v1[0], vals[2]) <= 1e-12 and relerr_norm(g1[0], grads[2]) <= 1e-10
    crit2 = gp.batched.BatchedCriterion(m, x, z, p, kind="ml" if kind == "zero" else "reml", max_bytes=4 << 20)
    v2, g2 = crit2.value_and_grad(TH)
    assert np.array_equal(v2, vals) and np.array_equal(g2, grads)


@pytest.mark.parametrize("n,d,p,kind,noise", [(257, 2, 0, "const", False), (130, 5, 4, "linear", False),
                                              (300, 3, 2, "const", True), (64, 2, 1, "zero", False),
                                              (513, 4, 10, "const", False)])
def test_batched_value_and_grad_variants(gp, n, d, p, kind, noise):
    """Ragged sizes, p = 0 / 4 / generic (10), a linear basis (q = d + 1 rows in the panels), the noisy kernel
    (gradient w.r.t. log tau2) and the zero-mean form: batched rows == the scalar value_and_grad."""
    x, z, _ = cases.data(n, d, 90 + n)
    th0 = cases.theta(d, 90 + n, noise=noise)
    N = 5
    TH = th0 + np.random.default_rng(n).uniform(-0.5, 0.5, size=(N, th0.shape[0]))
    m = _model(gp, kind, p, noise, th0)
    crit = gp.batched.BatchedCriterion(m, x, z, p, kind="ml" if kind == "zero" else "reml", noise=noise)
    vals, grads = crit.value_and_grad(TH)
    assert grads.shape == (N, th0.shape[0])
    f = m.negative_log_likelihood_zero_mean if kind == "zero" else m.negative_log_restricted_likelihood
    for i in range(N):
        v, gi = gp.num.value_and_grad(lambda t: f(t, x, z), TH[i])
        assert relerr(vals[i], float(v)) <= 1e-9, (i, vals[i], float(v))
        assert relerr_norm(grads[i], gi.cpu().numpy()) <= 1e-8, (i, grads[i], gi)


def test_batched_value_and_grad_not_pd(gp):
    """A particle whose matrix is not positive definite: +inf value, zero gradient row, neighbours untouched."""
    x, z, _ = cases.data(150, 2, 60)
    th = cases.theta(2, 60)
    TH = np.stack([th, th + np.array([-800.0, 0.0, 0.0]), th + 0.1])
    m = _model(gp, "const", 2, False, th)
    vals, grads = gp.batched.BatchedCriterion(m, x, z, 2).value_and_grad(TH)
    assert np.isinf(vals[1]) and np.all(grads[1] == 0.0)
    assert np.all(np.isfinite(vals[[0, 2]])) and np.all(np.isfinite(grads[[0, 2]]))


@pytest.mark.parametrize("case", [c[0] for c in _mb_cases()[0]])
def test_minibatch_criterion(gp, case):
    """MiniBatchCriterion against the reference's BatchDifferentiableSelectionCriterion (golden vectors):
    equal-size batches go through one batched launch sequence, the ragged last batch through the scalar path."""
    name, n, bs, d, p, kind, seed = next(c for c in _mb_cases()[0] if c[0] == case)
    g = _mb_golden()
    x, z, th = g[name + "/x"], g[name + "/z"], g[name + "/theta"]
    m = _model(gp, kind, p, False, th)
    loader = [(x[i:i + bs], z[i:i + bs]) for i in range(0, n, bs)]
    for red in ("mean", "sum"):
        c = gp.batched.MiniBatchCriterion(m, loader, p, kind="ml" if kind == "zero" else "reml", reduction=red)
        v = c.evaluate_pre_grad(th)
        assert isinstance(v, float) and relerr(v, float(g[f"{name}/{red}/value"])) <= TOL_LIK
        assert relerr_norm(np.asarray(c.gradient(th)), g[f"{name}/{red}/grad"]) <= TOL_LIK
        assert relerr(c.evaluate_no_grad(th), float(g[f"{name}/{red}/nograd"])) <= TOL_LIK
    # batches_per_eval: two batches per call, cycling
    c = gp.batched.MiniBatchCriterion(m, loader, p, kind="ml" if kind == "zero" else "reml", batches_per_eval=2)
    a, b = c.evaluate(th), c.evaluate(th)
    assert np.isfinite(a) and np.isfinite(b) and a != b
    with pytest.raises(RuntimeError):
        gp.batched.MiniBatchCriterion(m, loader, p).gradient(th)


def test_batched_grid_clients(gp):
    """Cross-sections and 2-D profiles of the criterion (modeldiagnosis/plotting.py:185-231, 300-327) through
    the batched sweep: same grids and the same values as the reference's per-point loop over the scalar path."""
    x, z, _ = cases.data(200, 2, 70)
    th = cases.theta(2, 70)
    m = _model(gp, "const", 2, False, th)
    crit = gp.batched.BatchedCriterion(m, x, z, 2)
    f = lambda t: m.negative_log_restricted_likelihood(t, x, z).item()
    cs = crit.cross_sections(th, n_points=7, delta=0.5)   # (a narrow box keeps K well conditioned)
    assert sorted(cs) == [0, 1, 2]
    for j, (grid, vals) in cs.items():
        assert np.allclose(grid, np.linspace(th[j] - 0.5, th[j] + 0.5, 7))
        for g, v in zip(grid[::3], vals[::3]):
            t = th.copy(); t[j] = g
            assert relerr(v, f(t)) <= 1e-9
    cs = crit.cross_sections(th, ind=[2], n_points=4, param_box=np.array([[0.0], [1.0]]))
    assert list(cs) == [2] and np.allclose(cs[2][0], np.linspace(0.0, 1.0, 4))
    p1, p2, vals = crit.profile_2d(th, (0, 1), n=5, factor=1.5)
    assert vals.shape == (5, 5) and np.isclose(p1[2], np.exp(th[0] / 2)) and np.isclose(p2[2], np.exp(-th[1]))
    t = th.copy(); t[0] = np.log(p1[4] ** 2); t[1] = np.log(1.0 / p2[1])
    assert relerr(vals[1, 4], f(t)) <= 1e-9


def _fisher_cases():
    from oracle.make_golden_fisher import FISHER_CASES
    return FISHER_CASES


@pytest.mark.parametrize("case", [c[0] for c in _fisher_cases()])
def test_fisher_information(gp, case):
    """Model.fisher_information / _cpd against the reference's core/fisher.py (golden vectors).  Both sides
    difference K at steps of 1e-3, so entry-level rounding of K (1e-13) is amplified by 1e3 cond(K): 1e-6."""
    import os
    name, n, d, p, kind, noise, seed = next(c for c in _fisher_cases() if c[0] == case)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fisher.npz"))
    x, th = g[name + "/x"], g[name + "/theta"]
    m = _model(gp, kind, p, noise, th)
    for form, fn in (("spd", m.fisher_information), ("cpd", m.fisher_information_cpd)):
        ref = g[name + "/" + form]
        got = fn(x, th).cpu().numpy()
        assert got.shape == ref.shape and np.allclose(got, got.T)
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) <= 1e-6, (form, got, ref)
    # default covparam = model.covparam
    assert np.array_equal(m.fisher_information(x).cpu().numpy(), m.fisher_information(x, th).cpu().numpy())


def test_edge_shapes_and_inputs(gp):
    """Edge cases the reference's call conventions allow: a (n,1) column for zi, NumPy covparam, a wide
    linear basis (q = d + 1 = 11), d = 32, n = 1, empty prediction sets."""
    rng = np.random.default_rng(11)
    # q = 11, d = 10, value + gradient against the oracle (analytic form checked against reference autograd)
    n, d = 260, 10
    x, z, xt = cases.data(n, d, 12, m=33)
    th = cases.theta(d, 12)
    m = _model(gp, "linear", 2, False, th)
    tp = torch.tensor(th, requires_grad=True)
    v = m.negative_log_restricted_likelihood(tp, x, z.reshape(-1, 1))
    (g,) = torch.autograd.grad(v, tp)
    P = np.hstack((np.ones((n, 1)), x))
    vr, gr = onp.reml_value_and_grad_analytic(x, z, P, 2, th)
    assert relerr(v.item(), vr) <= TOL_LIK and relerr_norm(g.numpy(), gr) <= TOL_LIK
    om = onp.OracleModel(cases.mean_fn("linear", np), lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, 2, cp, pw),
                         None, th, "linear_predictor")
    mu_r, var_r = onp.predict(om, x, z, xt)[:2]
    mu, var = m.predict(x, z, xt)
    s2 = float(np.exp(th[0]))
    assert np.max(np.abs(mu - mu_r)) / max(np.sqrt(s2), np.max(np.abs(mu_r))) <= 1e-8
    assert np.max(np.abs(var - var_r)) / s2 <= 1e-8
    mu0, var0 = m.predict(x, z, xt[:0])
    assert mu0.shape == (0,) and var0.shape == (0,)
    # d = 32 covariance
    x32 = rng.uniform(size=(70, 32))
    th32 = np.concatenate(([0.2], rng.normal(size=32) * 0.3 - 1.0))
    K = gp.kernel.maternp_covariance(gp.num.asarray(x32), None, 3, th32).cpu().numpy()
    assert relerr(K, onp.maternp_covariance(x32, x32, 3, th32)) <= TOL_COV
    with pytest.raises(gp._abi.GpmpError):
        gp.kernel.maternp_covariance(gp.num.asarray(rng.uniform(size=(5, 33))), None, 2, np.zeros(34))
    # n = 1 and n = 2
    for nn in (1, 2):
        xs, zs = rng.uniform(size=(nn, 2)), rng.normal(size=nn)
        ths = np.array([0.1, 0.3, -0.2])
        mz = _model(gp, "zero", 1, False, ths)
        ref = onp.negative_log_likelihood_zero_mean(
            onp.OracleModel(None, lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, 1, cp, pw), None, ths, "zero"),
            ths, xs, zs)
        assert relerr(mz.negative_log_likelihood_zero_mean(ths, xs, zs).item(), ref) <= TOL_LIK


def test_partitioned_factorisation_matches_local(gp):
    """The panel-partitioned evaluation (owner factors a column group, panel exchange, per-rank updates) run
    with a single rank must reproduce the local pipeline (to rounding: the local look-ahead path solves its
    panels by blocked substitution, the partitioned one by the tile inverse) and leave a state predict can use."""
    n, d = 1500, 5
    x, z, xt = cases.data(n, d, 31, m=40)
    th = cases.theta(d, 31)
    m = _model(gp, "linear", 2, False, th)
    v_local = m.negative_log_restricted_likelihood(th, x, z).item()
    v_dist, state = gp.dist.reml_value_distributed(m, th, x, z)
    assert relerr(v_dist, v_local) <= 1e-12
    ref = onp.negative_log_restricted_likelihood(
        onp.OracleModel(cases.mean_fn("linear", np), lambda a, b, cp, pw=False: onp.maternp_covariance(a, b, 2, cp, pw),
                        None, th, "linear_predictor"), th, x, z)
    assert relerr(v_dist, ref) <= TOL_LIK
    mean, var = gp.dist.predict_distributed(m, x, z, xt)
    mean1, var1 = m.predict(x, z, xt)
    assert np.array_equal(mean, mean1) and np.array_equal(var, var1)
    # the row-partitioned gradient (T^T rows, K^-1 rows, contraction of the owned rows) against autograd
    tp = torch.tensor(th, requires_grad=True)
    (g_local,) = torch.autograd.grad(m.negative_log_restricted_likelihood(tp, x, z), tp)
    v2, g_dist = gp.dist.reml_value_and_grad_distributed(m, th, x, z)
    assert relerr(v2, v_local) <= 1e-12 and relerr_norm(g_dist, g_local.numpy()) <= 1e-10
    m.covparam = th
    fitted = gp.dist.fit_distributed(m, x, z)
    mean2, var2 = gp.dist.predict_distributed(m, x, z, xt, fitted=fitted)
    # the partitioned path re-derives the tile inverses from the received factor (1 / L_ii instead of the
    # factorisation's own rsqrt), so predictions agree to rounding, not bit for bit
    assert relerr_norm(mean2, mean1) <= 1e-12 and relerr_norm(var2, var1) <= 1e-12


# ----------------------------------------------------------------------------------------------------
def test_headline_size_properties(gp):
    """n=8192, d=8 (BASELINE config 3): size-independent checks at full size.
    (1) the closed form d value / d log sigma2 = 0.5 ((n-q) - quad) must match the contracted gradient;
    (2) REML is invariant to adding a constant to z (the constant mean is profiled out);
    (3) gradient against a central finite difference of the value along a random direction."""
    x, z, th0 = cases.headline()
    m = _model(gp, "const", 2, False, th0)
    n = x.shape[0]
    xd, zd = gp.num.asarray(x), gp.num.asarray(z)
    tp = torch.tensor(th0, requires_grad=True)
    v = m.negative_log_restricted_likelihood(tp, xd, zd)
    (g,) = torch.autograd.grad(v, tp)
    quad = m.norm_k_sqrd(xd, zd, th0).item()
    assert abs(g[0].item() - 0.5 * ((n - 1) - quad)) <= 1e-8 * max(1.0, abs(g[0].item()))
    v_shift = m.negative_log_restricted_likelihood(th0, xd, zd + 3.0).item()
    assert abs(v_shift - v.item()) <= 1e-8 * abs(v.item())
    rng = np.random.default_rng(2)
    dirn = rng.standard_normal(th0.shape)
    dirn /= np.linalg.norm(dirn)
    h = 1e-4
    fp = m.negative_log_restricted_likelihood(th0 + h * dirn, xd, zd).item()
    fm = m.negative_log_restricted_likelihood(th0 - h * dirn, xd, zd).item()
    fd = (fp - fm) / (2 * h)
    assert abs(fd - float(g.numpy() @ dirn)) <= 1e-5 * max(1.0, abs(fd))


@pytest.mark.parametrize("n", [4096, 16384])
def test_early_inverse_sizes_properties(gp, n):
    """Sizes at which the leading block of T = L^-1 is computed under the tail of the factorisation (potrf.cu
    early_inverse: n a multiple of the group width with at least 8 groups, up to 16384): the gradient that consumes
    that block must satisfy the closed form d value / d log sigma2 = 0.5 ((n-q) - quad) and a central finite
    difference of the value (which runs WITHOUT the early block: value-sized workspace) along a random direction."""
    x, z, th0 = cases.headline(n=n)
    m = _model(gp, "const", 2, False, th0)
    xd, zd = gp.num.asarray(x), gp.num.asarray(z)
    tp = torch.tensor(th0, requires_grad=True)
    v = m.negative_log_restricted_likelihood(tp, xd, zd)
    (g,) = torch.autograd.grad(v, tp)
    quad = m.norm_k_sqrd(xd, zd, th0).item()
    closed = 0.5 * ((n - 1) - quad)
    err = abs(g[0].item() - closed) / max(1.0, abs(closed))
    print(f"[parity] n={n}: d/dlog sigma2 {g[0].item():.12g} vs closed form {closed:.12g} (rel {err:.2e})")
    assert err <= 1e-8
    rng = np.random.default_rng(3)
    dirn = rng.standard_normal(th0.shape)
    dirn /= np.linalg.norm(dirn)
    h = 1e-4
    fp = m.negative_log_restricted_likelihood(th0 + h * dirn, xd, zd).item()
    fm = m.negative_log_restricted_likelihood(th0 - h * dirn, xd, zd).item()
    fd = (fp - fm) / (2 * h)
    assert abs(fd - float(g.numpy() @ dirn)) <= 1e-5 * max(1.0, abs(fd))
    # the value computed with and without the gradient-sized workspace is the same number
    assert abs(v.item() - m.negative_log_restricted_likelihood(th0, xd, zd).item()) <= 1e-12 * abs(v.item())
