"""gpmp.core.Model for the exact-GP inner loop on the B200.

Same constructor, method names, argument order and return conventions as gpmp/core/model.py:22-696 for
the methods on the hot path; the arithmetic of gpmp/core/likelihood.py, linalg.py, kriging.py and
sample_paths.py is replaced by the device pipelines of libgpmp_b200.so:

  likelihoods   K -> blocked DMMA Cholesky with z and the mean basis whitened as extra rows -> CGS2 of the
                whitened rows (REML without the n x (n-q) contrast matrix of core/linalg.py:49-88)
  gradients     T = L^-1, K^-1 = T^T T, M = K^-1 - U^T U, contraction against regenerated dK tiles
  predict       V = K(xt, xi) L^-T chunk by chunk, row reductions against [Q~; r]; lambda only on request

Two ways in, chosen per call from what the user's covariance callable returns:
  fused        the callable is a plain gp.kernel.maternp_covariance(x, y, p, param): K is built by the
               Matern tile kernel inside the pipeline and the gradient never materialises dK;
  composable   anything else: the callable's own (device) ops build K, the likelihood op returns
               dvalue/dK = M/2 and autograd continues through the callable.
"""
from __future__ import annotations

import math
import warnings

import numpy as np
import torch

from . import _abi, kernel, num, ops


def _validate_model_mean(meantype, mean, meanparam):
    # gpmp/core/utils.py:84-119
    if meantype not in {"zero", "parameterized", "linear_predictor"}:
        raise ValueError("meantype must be one of 'zero', 'parameterized', or 'linear_predictor'")
    if meantype == "zero" and mean is not None:
        raise ValueError("For meantype 'zero', mean must be None")
    if meantype in ("parameterized", "linear_predictor") and not callable(mean):
        raise TypeError("For meantype 'parameterized' or 'linear_predictor', mean must be a callable function")


def _ensure_shapes_and_type(xi=None, zi=None, xt=None, convert=True):
    # gpmp/core/utils.py:19-81
    if xi is not None:
        assert len(xi.shape) == 2, "xi should be a 2D array"
    if zi is not None:
        if len(zi.shape) == 2:
            assert zi.shape[1] == 1, "zi should only have one column if it's a 2D array"
            zi = zi.reshape(-1)
        else:
            assert len(zi.shape) == 1, "zi should be 1D or a 2D column array"
    if xt is not None:
        assert len(xt.shape) == 2, "xt should be a 2D array"
    if xi is not None and zi is not None:
        assert xi.shape[0] == zi.shape[0], "xi and zi must have the same number of rows"
    if xi is not None and xt is not None:
        assert xi.shape[1] == xt.shape[1], "xi and xt must have the same number of columns"
    # device placement is not optional here (convert=False only skips nothing: data must be on the GPU)
    xi = ops.to_device(xi)
    zi = ops.to_device(zi)
    xt = ops.to_device(xt)
    return xi, zi, xt


def _as_param(p):
    if p is None:
        return None
    return num.asparam(p)


class Model:
    """GP model: mean + covariance callables and their parameters (gpmp/core/model.py:136-166)."""

    def __init__(self, mean, covariance, meanparam=None, covparam=None, meantype="linear_predictor"):
        _validate_model_mean(meantype, mean, meanparam)
        self.meantype = meantype
        self.mean = mean
        self.meanparam = meanparam
        self.covparam = covparam
        self.covariance = covariance
        # True: the user's `mean` callable is written against a HOST array namespace (GPmp's own gnp when this
        # library is bound into it, gpmp_b200/dropin.py); it then receives host copies of the points and its
        # result is moved to the device here
        self.host_mean = False

    def mean_values(self, x, meanparam=None):
        """mean(x, meanparam) on the device, whatever array namespace the callable is written against."""
        xin = x.cpu() if (self.host_mean and torch.is_tensor(x) and x.device.type != "cpu") else x
        return ops.to_device(self.mean(xin, self.meanparam if meanparam is None else meanparam))

    def __repr__(self):
        return "<gpmp_b200.core.Model object> " + hex(id(self))

    # ------------------------------------------------------------------ helpers
    def _same_set_cov(self, x, covparam):
        """K(x, x) from the user's callable, lazily when it is a plain Matern-p covariance."""
        with kernel.capture():
            K = self.covariance(x, x, covparam)
        fused = isinstance(K, kernel.LazyMatern) and K.y is None and K.x is x
        return K, fused

    def _criterion(self, covparam, xi, zi, P):
        covparam = _as_param(covparam)
        K, fused = self._same_set_cov(xi, covparam)
        try:
            if fused:
                return ops.fused_likelihood(K.param, zi, xi, P, K.p)
            return ops.likelihood_from_K(kernel.materialize(K), zi, P)
        except torch.linalg.LinAlgError:
            return num.safe_inf()

    def _basis(self, x, meanparam=None):
        P = self.mean_values(x, meanparam)
        if P.dim() == 1:
            P = P.reshape(-1, 1)
        return P

    # ------------------------------------------------------------------ likelihoods
    def negative_log_likelihood_zero_mean(self, covparam, xi, zi):
        """0.5 (n log 2pi + log det K + z^T K^-1 z)  (core/likelihood.py:18-52); +inf when K is not PD."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        v = self._criterion(covparam, xi, zi, None)
        return v.reshape(()) if torch.isfinite(v) else v

    def negative_log_likelihood(self, meanparam, covparam, xi, zi):
        """Zero-mean NLL of z - mean(x, meanparam) (core/likelihood.py:55-89); differentiable in both."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        prior_mean = self.mean_values(xi, _as_param(meanparam)).reshape(-1)
        v = self._criterion(covparam, xi, zi - prior_mean, None)
        return v.reshape(()) if torch.isfinite(v) else v

    def negative_log_restricted_likelihood(self, covparam, xi, zi):
        """0.5 ((n-q) log 2pi + log det(W^T K W) + (W^T z)^T (W^T K W)^-1 W^T z)  (core/likelihood.py:92-129),
        evaluated through the Cholesky-whitened mean basis instead of the contrast matrix W."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        P = self._basis(xi)
        v = self._criterion(covparam, xi, zi, P)
        return v.reshape(()) if torch.isfinite(v) else v

    # ------------------------------------------------------------------ secondary norms (core/linalg.py:113-141)
    def norm_k_sqrd_with_zero_mean(self, xi, zi, covparam):
        """z^T K^-1 z."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        state, out = self._fit(xi, zi, None, _as_param(covparam))
        return torch.tensor(ops.read_small(out)[2], dtype=torch.float64)

    def norm_k_sqrd(self, xi, zi, covparam):
        """(W^T z)^T (W^T K W)^-1 (W^T z) for the linear-predictor mean."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        state, out = self._fit(xi, zi, self._basis(xi), _as_param(covparam))
        return torch.tensor(ops.read_small(out)[2], dtype=torch.float64)

    def k_inverses(self, xi, zi, covparam):
        """(z^T K^-1 z, K^-1 1, K^-1 z)  (core/linalg.py:121-129), by one factorisation with z and the ones
        vector carried as right-hand-side rows (no explicit inverse)."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        with torch.no_grad():
            K = kernel.materialize(self.covariance(xi, xi, _as_param(covparam)))
            rhs = torch.stack((zi, torch.ones_like(zi)))
            fac = ops.potrf(K, extra_rows=rhs)
            rows = fac.A[fac.n:, :]
            ops.trsm_rows(fac, rows, trans=1)
            kinv_z, kinv_1 = rows[0, : fac.n].clone(), rows[1, : fac.n].clone()
            return torch.dot(zi, kinv_z), kinv_1, kinv_z

    # ------------------------------------------------------------------ prediction
    def _fit(self, xi, zi, P, covparam, with_inverse=False):
        """Factor K(xi, xi) with zi (and P) whitened along -> (FitState, out_dev)."""
        K, fused = self._same_set_cov(xi, covparam)
        if fused:
            vals = ops.host_values(K.param)
            spec = ops._spec_from_param(K.p, xi.shape[1], vals)
            state, out = ops.lik_value(spec, None, xi, zi, P, with_inverse)
            state.kernel = (K.p, K.param)
        else:
            Kd = kernel.materialize(K)
            Kd = Kd.detach()
            Kd = Kd if Kd.stride(1) == 1 else Kd.contiguous()
            state, out = ops.lik_value(None, Kd, None, zi, P, with_inverse)
            state.kernel = None
        return state, out

    # ------------------------------------------------------------------ Fisher information
    def fisher_information(self, xi, covparam=None, epsilon=1e-3):
        """0.5 tr(K^-1 dK_i K^-1 dK_j), dK by 5-point differences (core/model.py:509-538, core/fisher.py:18-78)."""
        from . import fisher
        return fisher.fisher_information(self, xi, covparam=covparam, epsilon=epsilon)

    def fisher_information_cpd(self, xi, covparam=None, epsilon=1e-3):
        """Contrast-space form for a linear-predictor mean (core/model.py:540-575, core/fisher.py:81-155)."""
        from . import fisher
        return fisher.fisher_information_cpd(self, xi, covparam=covparam, epsilon=epsilon)

    # ------------------------------------------------------------------ leave-one-out
    def loo(self, xi, zi, convert_in=True, convert_out=False):
        """Leave-one-out predictions, variances and errors by virtual cross-validation
        (core/model.py:309-343, core/loo.py:65-130), from the K^-1 / Pi the gradient pipeline builds."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi, convert=convert_in)
        covparam = _as_param(self.covparam)
        P = None
        zc, prior = zi, None
        if self.meantype == "linear_predictor":
            P = self._basis(xi)
        elif self.meantype == "parameterized":
            prior = self.mean_values(xi, _as_param(self.meanparam)).reshape(-1)
            zc = zi - prior
        with torch.no_grad():
            state, out = self._fit(xi, zc, P, covparam, with_inverse=True)
            if ops.read_small(out)[6] != 0.0:
                raise torch.linalg.LinAlgError("loo: K(xi, xi) is not positive-definite")
            zloo, s2, e = ops.lik_loo(state, zc.contiguous())
            if prior is not None:
                zloo = zloo + prior
        if convert_out:
            return num.to_np(zloo), num.to_np(s2), num.to_np(e)
        return zloo, s2, e

    def fit(self, xi, zi, state=None):
        """Factor K(xi, xi) once (zi and the mean basis whitened along) and return a `Fitted` handle whose
        predict / conditional_sample_paths_chunked reuse the factor.  `state` lets a caller supply a fitted
        state computed elsewhere (gpmp_b200.dist.fit_distributed: factorisation partitioned over GPUs)."""
        xi, zi, _ = _ensure_shapes_and_type(xi=xi, zi=zi)
        covparam = _as_param(self.covparam)
        P = None
        zc, zi_prior = zi, None
        if self.meantype == "linear_predictor":
            P = self._basis(xi)
        elif self.meantype == "parameterized":
            if self.meanparam is None:
                raise ValueError("For meantype 'parameterized', meanparam should not be None.")
            zi_prior = self.mean_values(xi, _as_param(self.meanparam)).reshape(-1)
            zc = zi - zi_prior
        elif self.meantype != "zero":
            raise ValueError(f"Invalid meantype {self.meantype}.")
        if state is None:
            with torch.no_grad():
                state, out = self._fit(xi, zc, P, covparam)
                if ops.read_small(out)[6] != 0.0:
                    raise torch.linalg.LinAlgError("fit: K(xi, xi) is not positive-definite")
        elif state.spec is not None:
            state.kernel = (state.spec.p, covparam)  # lets the chunks recognise the plain Matern cross-covariance
        return Fitted(self, state, xi, zc, covparam)

    def predict(self, xi, zi, xt, return_lambdas=False, zero_neg_variances=True, convert_in=True,
                convert_out=True):
        """Posterior mean / variance at xt given (xi, zi)  (core/model.py:227-307).

        "zero": simple kriging; "linear_predictor": universal kriging with basis mean(x, .);
        "parameterized": simple kriging of zi - mean(xi, meanparam), prior mean added back.
        Returns (mean[nt], var[nt]) as NumPy arrays (convert_out) or device tensors, plus the kriging
        weights lambda_t (ni x nt, device) when return_lambdas.
        """
        return self.fit(xi, zi).predict(xt, return_lambdas=return_lambdas, zero_neg_variances=zero_neg_variances,
                                        convert_out=convert_out)

    # ------------------------------------------------------------------ kriging predictors (core/model.py:196-222)
    def kriging_predictor_with_zero_mean(self, xi, xt, return_type=0):
        """(lambda_t, posterior variance) of simple kriging (core/kriging.py:35-67); return_type -1: no variance,
        0: marginal variances (m,), 1: the full posterior covariance (m, m).  Device tensors."""
        return self._kriging(xi, xt, None, return_type)

    def kriging_predictor(self, xi, xt, return_type=0):
        """(lambda_t, posterior variance) of universal kriging with the basis mean(x, .) (core/kriging.py:70-116,
        by block elimination instead of the (n+q) saddle system)."""
        xi_ = ops.to_device(xi)
        return self._kriging(xi_, xt, self._basis(xi_), return_type)

    def _kriging(self, xi, xt, P, return_type):
        if return_type not in (-1, 0, 1):
            raise ValueError("return_type must be in {-1, 0, 1}")
        xi_, _, xt_ = _ensure_shapes_and_type(xi=xi, xt=xt)
        covparam = _as_param(self.covparam)
        with torch.no_grad():
            zero = torch.zeros(xi_.shape[0], dtype=torch.float64, device=xi_.device)
            state, out = self._fit(xi_, zero, P, covparam)
            if ops.read_small(out)[6] != 0.0:
                raise torch.linalg.LinAlgError("kriging_predictor: K(xi, xi) is not positive-definite")
        fitted = Fitted(self, state, xi_, zero, covparam)
        return fitted.weights_and_variance(xt_, return_type)

    # ------------------------------------------------------------------ sample paths
    def sample_paths(self, xt, nb_paths, method="chol", check_result=True):
        """nb_paths draws of GP(0, k) at xt: C @ N(0, I) (core/sample_paths.py:18-63).
        method='chol': K(xt, xt) = C C^T by the device Cholesky; check_result=True raises when K is not positive
        definite (the reference's NaN check, :46-49), False skips the device->host read of the status word.
        method='svd' (:50-58, for covariance matrices Cholesky cannot factor): C = U sqrt(s) U^T, the symmetric
        square root of K -- computed here without an eigensolver, by the coupled Newton-Schulz iteration on the
        DMMA GEMM (see _symmetric_sqrt)."""
        if method not in ("chol", "svd"):
            raise ValueError("method must be 'chol' or 'svd'")
        xt_ = ops.to_device(xt)
        normals = torch.randn(xt_.shape[0], nb_paths, dtype=torch.float64, device=xt_.device,
                              generator=_generator())
        return self.sample_paths_from_normals(xt_, normals, check_result=check_result, method=method)

    def sample_paths_from_normals(self, xt, normals, check_result=True, method="chol"):
        """Deterministic part of sample_paths: C @ normals (the map parity is defined on)."""
        if method not in ("chol", "svd"):
            raise ValueError("method must be 'chol' or 'svd'")
        xt_ = ops.to_device(xt)
        normals = ops.to_device(normals)
        with torch.no_grad():
            K = kernel.materialize(self.covariance(xt_, xt_, _as_param(self.covparam)))
            Nt = ops.transpose(normals)  # paths x nt
            if method == "svd":
                Croot = _symmetric_sqrt(K)  # symmetric: C @ normals = (normals^T C)^T
                return ops.transpose(ops.gemm_nt(Nt, Croot)).contiguous()
            fac = ops.potrf(K, check_pd=bool(check_result))  # LinAlgError when not PD and checked
            nt = fac.n
            L = fac.A[:nt, :nt]
            return ops.gemm_nt(L, Nt, tri=_abi.TRI_A_LOWER).contiguous()

    def conditional_sample_paths(self, ztsim, xi_ind, zi, xt_ind, lambda_t, convert_out=True):
        """Conditioning by kriging (core/sample_paths.py:66-119):
        ztsim[xt_ind] + lambda_t^T (zi - ztsim[xi_ind])."""
        return self._condition(ztsim, xi_ind, ops.to_device(zi).reshape(-1), xt_ind, lambda_t, None, convert_out)

    def conditional_sample_paths_parameterized_mean(self, ztsim, xi, xi_ind, zi, xt, xt_ind, lambda_t,
                                                    convert_out=True):
        """Same with a parameterised prior mean (core/sample_paths.py:122-182)."""
        xi_, zi_, xt_ = _ensure_shapes_and_type(xi=xi, zi=zi, xt=xt)
        mp = _as_param(self.meanparam)
        zc = zi_ - self.mean_values(xi_, mp).reshape(-1)
        zt_prior = self.mean_values(xt_, mp).reshape(-1, 1)
        return self._condition(ztsim, xi_ind, zc, xt_ind, lambda_t, zt_prior, convert_out)

    def _condition(self, ztsim, xi_ind, zc, xt_ind, lambda_t, zt_prior, convert_out):
        with torch.no_grad():
            zs = ops.to_device(ztsim)
            dev = zs.device
            xi_ind = torch.as_tensor(np.asarray(xi_ind).reshape(-1), dtype=torch.long, device=dev)
            xt_ind = torch.as_tensor(np.asarray(xt_ind).reshape(-1), dtype=torch.long, device=dev)
            delta = zc.reshape(-1, 1) - zs[xi_ind, :]          # ni x paths
            out = ops.padded(zs[xt_ind, :])                    # nt x paths (accumulated in place)
            lam = ops.to_device(lambda_t, requires_contiguous=False)  # ni x nt
            # rows lambda_t^T (nt x ni): free when lambda_t is the transposed view predict() returns
            lamT = lam.t()
            if lamT.stride(1) != 1 or lamT.stride(0) % 2 or lamT.data_ptr() % 16:
                lamT = ops.transpose(lam)
            ops.gemm_nt(lamT, ops.transpose(delta), C_out=out, alpha=1.0, beta=1.0)
            if zt_prior is not None:
                out = out + zt_prior
            out = out.contiguous()
        return num.to_np(out) if convert_out else out


class Fitted:
    """A model fitted to (xi, zi): the device factor L (both ways), the whitened data, Q~ / R~ of the whitened
    mean basis.  Everything that follows the factorisation -- prediction, kriging weights, conditioning of
    sample paths -- streams test points through it chunk by chunk; lambda_t is formed only on request."""

    def __init__(self, model, state, xi, zc, covparam):
        self.model, self.state, self.xi, self.zc, self.covparam = model, state, xi, zc, covparam
        self.n = xi.shape[0]
        self.ld = ops._round_ld(self.n)
        self.chunk = int(max(128, min(32768, (1 << 31) // (8 * self.ld))))

    def _chunk(self, xtc, Pt, ktt, Vt, mode, return_dots=False):
        """One chunk through gpmp_predict_chunk; takes the callable's own cross-covariance when it is not the
        plain Matern of the same-set call."""
        model, state, n = self.model, self.state, self.n
        if state.spec is not None:
            with kernel.capture():
                Kx = model.covariance(self.xi, xtc, self.covparam)
            same = (isinstance(Kx, kernel.LazyMatern) and Kx.y is xtc and Kx.x is self.xi
                    and getattr(state, "kernel", None) is not None and Kx.p == state.kernel[0]
                    and Kx.param is state.kernel[1])
            if same:
                return ops.predict_chunk(state, xtc, Pt, ktt, Vt, mode, return_dots)
            Vt[:, :n].copy_(kernel.materialize(Kx).t())
            saved, state.spec = state.spec, None
            try:
                return ops.predict_chunk(state, xtc, Pt, ktt, Vt, mode, return_dots)
            finally:
                state.spec = saved
        Kx = kernel.materialize(model.covariance(self.xi, xtc, self.covparam))
        Vt[:, :n].copy_(Kx.t())
        return ops.predict_chunk(state, xtc, Pt, ktt, Vt, mode, return_dots)

    def _basis_and_prior(self, xt):
        model = self.model
        Pt = model._basis(xt).contiguous() if model.meantype == "linear_predictor" else None
        prior = None
        if model.meantype == "parameterized":
            prior = model.mean_values(xt, _as_param(model.meanparam)).reshape(-1)
        return Pt, prior

    def predict(self, xt, return_lambdas=False, zero_neg_variances=True, convert_out=True):
        xt = ops.to_device(xt)
        assert xt.dim() == 2 and xt.shape[1] == self.xi.shape[1], "xi and xt must have the same number of columns"
        n, m = self.n, xt.shape[0]
        with torch.no_grad():
            Pt, prior = self._basis_and_prior(xt)
            ktt = ops.to_device(self.model.covariance(xt, None, self.covparam, pairwise=True)).reshape(-1).contiguous()
            lam_rows = ops._empty((m, self.ld)) if return_lambdas else None
            mean, var = ops._empty((m,)), ops._empty((m,))
            for c0 in range(0, m, self.chunk):
                c1 = min(m, c0 + self.chunk)
                Vt = lam_rows[c0:c1] if return_lambdas else ops._empty((c1 - c0, self.ld))
                mu, s2 = self._chunk(xt[c0:c1], None if Pt is None else Pt[c0:c1], ktt[c0:c1], Vt,
                                     1 if return_lambdas else 0)
                mean[c0:c1] = mu
                var[c0:c1] = s2
            if prior is not None:
                mean = mean + prior
            if bool((var < 0.0).any()):
                warnings.warn("Negative variances detected. Consider using jitter.", RuntimeWarning)
            if zero_neg_variances:
                var = torch.clamp_min(var, 0.0)
        if convert_out:
            mean, var = num.to_np(mean), num.to_np(var)
        if return_lambdas:
            return mean, var, lam_rows[:, :n].t()
        return mean, var

    def weights_and_variance(self, xt, return_type=0):
        """(lambda_t (n x m), posterior variance) like core/kriging.py's predictors: return_type -1 -> None,
        0 -> marginal variances (m,), 1 -> posterior covariance matrix (m, m) = K_tt - V V^T + E E^T, where V rows are
        the whitened cross-covariances and E rows the mean-basis corrections e_t (the reference
        forms K_tt - [lambda; mu]^T [K_it; P_t^T], core/kriging.py:192-197).  Device tensors, variances not clamped."""
        xt = ops.to_device(xt)
        n, m, q = self.n, xt.shape[0], self.state.q
        use_basis = q > 0
        with torch.no_grad():
            Pt = self.model._basis(xt).contiguous() if use_basis else None
            ktt = ops.to_device(self.model.covariance(xt, None, self.covparam, pairwise=True)).reshape(-1).contiguous()
            cov = None
            if return_type == 1:
                # pass 1: V (mode 0) and the e_t records of every chunk
                V = ops._empty((m, self.ld))
                E = ops._empty((m, ops._round_ld(max(q, 1)))) if use_basis else None
                for c0 in range(0, m, self.chunk):
                    c1 = min(m, c0 + self.chunk)
                    _, _, dots = self._chunk(xt[c0:c1], None if Pt is None else Pt[c0:c1], ktt[c0:c1], V[c0:c1], 0,
                                             return_dots=True)
                    if use_basis:
                        E[c0:c1, :q].copy_(dots[:, :q])
                Ktt = kernel.materialize(self.model.covariance(xt, xt, self.covparam))
                cov = ops.padded(Ktt)
                ops.gemm_nt(V[:, :n], V[:, :n], C_out=cov, alpha=-1.0, beta=1.0)
                if use_basis:
                    ops.gemm_nt(E[:, :q], E[:, :q], C_out=cov, alpha=1.0, beta=1.0)
                del V
            lam_rows = ops._empty((m, self.ld))
            var = ops._empty((m,))
            for c0 in range(0, m, self.chunk):
                c1 = min(m, c0 + self.chunk)
                _, s2 = self._chunk(xt[c0:c1], None if Pt is None else Pt[c0:c1], ktt[c0:c1], lam_rows[c0:c1], 1)
                var[c0:c1] = s2
        lam = lam_rows[:, :n].t()
        if return_type == -1:
            return lam, None
        return lam, (var if return_type == 0 else cov)

    def conditional_sample_paths_chunked(self, ztsim, xi_ind, xt, xt_ind, convert_out=True):
        """Conditioning by kriging without ever forming lambda_t (the reference's
        core/sample_paths.py:66-182 needs the ni x nt weights, 262 GB at n=32768, nt=1e6):
            ztsimc = ztsim[xt_ind] + W delta~^T (+ prior mean),   W rows = v_t - Q~ e_t,  delta~ = L^-1 delta
        where delta = zi (centred) - ztsim[xi_ind].  One predict chunk + one DMMA GEMM per chunk of xt."""
        xt = ops.to_device(xt)
        with torch.no_grad():
            zs = ops.to_device(ztsim)
            dev = zs.device
            xi_ind = torch.as_tensor(np.asarray(xi_ind).reshape(-1), dtype=torch.long, device=dev)
            xt_ind = torch.as_tensor(np.asarray(xt_ind).reshape(-1), dtype=torch.long, device=dev)
            n, m, npaths = self.n, xt.shape[0], zs.shape[1]
            delta_rows = ops.transpose(self.zc.reshape(-1, 1) - zs[xi_ind, :])  # paths x n
            ops.lik_trsm_rows(self.state, delta_rows, trans=0)
            Pt, prior = self._basis_and_prior(xt)
            ktt = ops.to_device(self.model.covariance(xt, None, self.covparam, pairwise=True)).reshape(-1).contiguous()
            out = ops._empty((m, ops._round_ld(npaths)))[:, :npaths]
            for c0 in range(0, m, self.chunk):
                c1 = min(m, c0 + self.chunk)
                Vt = ops._empty((c1 - c0, self.ld))
                self._chunk(xt[c0:c1], None if Pt is None else Pt[c0:c1], ktt[c0:c1], Vt, 2)
                oc = out[c0:c1]
                oc.copy_(zs[xt_ind[c0:c1], :])
                ops.gemm_nt(Vt[:, :n], delta_rows, C_out=oc, alpha=1.0, beta=1.0)
            if prior is not None:
                out = out + prior.reshape(-1, 1)
            out = out.contiguous()
        return num.to_np(out) if convert_out else out


_gen = None


def _generator():
    """Backend-global generator seeded 1234, like gpmp/num/torch_backend.py:894-915 (draws themselves are
    device draws: bit-parity with the CPU generator is not defined)."""
    global _gen
    if _gen is None:
        _gen = torch.Generator(device=ops.device())
        _gen.manual_seed(1234)
    return _gen


def set_seed(seed):
    _generator().manual_seed(int(seed))


def _symmetric_sqrt(K, max_iter=120):
    """Symmetric (principal) square root of a symmetric positive SEMI-definite matrix on the device -- what the
    reference's U sqrt(s) V^T of gnp.svd(K, hermitian=True) is (core/sample_paths.py:50-55) -- by the coupled
    Newton-Schulz iteration, which needs nothing but products (three DMMA GEMMs per step):

        Y0 = K / ||K||_F,  Z0 = I,   W = 3 I - Z Y,   Y <- Y W / 2,   Z <- W Z / 2,      sqrt(K) = sqrt(||K||_F) lim Y.

    The ORDER of the factors matters for stability (Y W and W Z, never Y W^T or W Z^T: in floating point the iterates
    are only nearly symmetric, and the transposed forms amplify that defect), so the second operand of every NT
    product is transposed explicitly first (n^2 against the product's n^3).
    Eigenvalues of K / ||K||_F lie in [0, 1]: the iteration converges for all of them (zero stays zero); a direction
    of relative size lambda needs about log_1.5(lambda^-1/2) steps, after which convergence is quadratic.  Z tends
    to the INVERSE root, which does not exist for a singular K, so rounding noise in the null directions eventually
    grows: the step with the smallest change of Y is kept (on singular test matrices -- duplicated points, dense
    1-d designs with cond ~ 5e17: C C^T = K to 2e-8 ||K||, paths within 1e-7 of those of the SVD root, the size of the
    sqrt(eps) noise both carry in the null space; well-conditioned matrices converge to rounding)."""
    n = K.shape[0]
    scale = float(torch.linalg.norm(K))
    if not (scale > 0.0):
        return K.clone()
    eye3 = 3.0 * torch.eye(n, dtype=torch.float64, device=K.device)
    Y = ops.padded(K / scale)
    Z = ops.padded(torch.eye(n, dtype=torch.float64, device=K.device))
    best, best_Y, best_it = float("inf"), Y, 0
    for it in range(max_iter):
        W = ops.padded(eye3)
        ops.gemm_nt(Z, ops.transpose(Y), C_out=W, alpha=-1.0, beta=1.0)   # W = 3 I - Z Y
        Yn = ops.gemm_nt(Y, ops.transpose(W), alpha=0.5)                   # Y W / 2
        Z = ops.gemm_nt(W, ops.transpose(Z), alpha=0.5)                    # W Z / 2
        change = float(torch.linalg.norm(Yn - Y) / torch.linalg.norm(Yn))
        Y = Yn
        if change < best:
            best, best_Y, best_it = change, Y, it
        if change < 1e-15 or (change > 4.0 * best and it > best_it + 1):
            break
    return ops.padded(math.sqrt(scale) * best_Y)
