"""Particle-batched criterion evaluation: the replacement for the one-theta-at-a-time loop of
gpmp/mcmc/param_posterior.py:739-759 (`logpdf_temp`), also usable by multi-start restarts and
criterion grids.

    crit = BatchedCriterion(model, xi, zi, p=2, kind="reml")
    vals = crit(thetas)                        # (N, 1+d) -> (N,) criterion values, +inf where K is not PD
    logp = crit.logpdf_temp(thetas, T, box)    # -vals / T, -inf outside the sampling box

Every row is an independent REML / ML evaluation on the same (xi, zi); rows are block-partitioned over
the ranks of a torch.distributed group when one is given (no data-path collective except the final
all-gather of N float64 values).
"""
from __future__ import annotations

import numpy as np
import torch

from . import dist as gdist
from . import ops


def _mean_values(model, x):
    """mean(x, meanparam) on the device for gpmp_b200.core.Model (which knows whether the callable wants host
    points) or any object with the reference Model's `mean` / `meanparam` attributes."""
    f = getattr(model, "mean_values", None)
    return f(x) if f is not None else ops.to_device(model.mean(x, model.meanparam))


class BatchedCriterion:
    def __init__(self, model, xi, zi, p, kind="reml", noise=False, group=None, max_bytes=None):
        if kind not in ("reml", "ml"):
            raise ValueError("kind must be 'reml' or 'ml'")
        self.x = ops.to_device(xi)
        z = ops.to_device(zi).reshape(-1)
        self.p, self.noise, self.kind = int(p), bool(noise), kind
        self.P = None
        if kind == "reml":
            P = _mean_values(model, self.x)
            self.P = (P.reshape(-1, 1) if P.dim() == 1 else P).contiguous()
        elif model.meantype == "parameterized":
            z = z - _mean_values(model, self.x).reshape(-1)
        self.z = z.contiguous()
        self.group = group
        self.max_bytes = max_bytes
        self._work = None  # device workspace, allocated at the first sweep and reused
        self._gwork, self._grows = None, 0

    def values_device(self, thetas):
        """thetas: (N, 1+noise+d) array-like -> (values, info) device tensors for the local shard."""
        th = ops.to_device(thetas)
        n = self.x.shape[0]
        q = 0 if self.P is None else self.P.shape[1]
        if self._work is None or self._rows < th.shape[0]:
            self._work = ops.criterion_batched_workspace(n, q, th.shape[0], self.max_bytes)
            self._rows = th.shape[0]
        return ops.criterion_batched(th, self.x, self.z, self.P, self.p, self.noise, self.max_bytes, self._work)

    def value_and_grad(self, thetas, convert_out=True):
        """Criterion values and gradients at every row of thetas in one launch sequence: the batched form of
        `gnp.value_and_grad(criterion)` per particle (mcmc/svgd.py:310-313, torch_backend.py:516-533).
        Returns (values[N], grads[N, 1+noise+d]); rows are block-partitioned over ranks like the values."""
        th = np.asarray(thetas.detach().cpu() if torch.is_tensor(thetas) else thetas, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(1, -1)
        N = th.shape[0]
        lo, hi = gdist.shard_bounds(N, self.group)
        thd = ops.to_device(th[lo:hi])
        n, d = self.x.shape
        q = 0 if self.P is None else self.P.shape[1]
        if self._gwork is None or self._grows < thd.shape[0]:
            self._gwork = ops.criterion_batched_grad_workspace(n, q, d, thd.shape[0], self.max_bytes)
            self._grows = thd.shape[0]
        vals, grads, _ = ops.criterion_batched_grad(thd, self.x, self.z, self.P, self.p, self.noise,
                                                    self.max_bytes, self._gwork)
        vals = gdist.all_gather_rows(vals, N, self.group)
        grads = gdist.all_gather_rows(grads, N, self.group)
        if convert_out:
            return vals.cpu().numpy(), grads.cpu().numpy()
        return vals, grads

    def __call__(self, thetas, convert_out=True):
        th = np.asarray(thetas.detach().cpu() if torch.is_tensor(thetas) else thetas, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(1, -1)
        N = th.shape[0]
        lo, hi = gdist.shard_bounds(N, self.group)
        vals, _ = self.values_device(th[lo:hi])
        vals = gdist.all_gather_rows(vals, N, self.group)
        return vals.cpu().numpy() if convert_out else vals

    def logpdf_temp(self, thetas, temperature, lower=None, upper=None):
        """-criterion / T with -inf outside [lower, upper] (param_posterior.py:739-759)."""
        th = np.asarray(thetas.detach().cpu() if torch.is_tensor(thetas) else thetas, dtype=np.float64)
        if th.ndim == 1:
            th = th.reshape(1, -1)
        out = -self(th) / float(temperature)
        out = np.where(np.isnan(out), -np.inf, out)
        if lower is not None and upper is not None:
            inside = np.all((th >= np.asarray(lower)) & (th <= np.asarray(upper)), axis=1)
            out = np.where(inside, out, -np.inf)
        return out


    # ------------------------------------------------------------------ grid clients (model diagnosis)
    def cross_sections(self, param_opt, ind=None, n_points=100, delta=5.0, param_box=None):
        """Criterion along each coordinate through `param_opt` -- the data of
        `plot_selection_criterion_crosssections` (gpmp/modeldiagnosis/plotting.py:72-231, same defaults):
        for parameter ind[a] the grid runs over [opt - delta, opt + delta] or over
        [param_box[0, a], param_box[1, a]].  All len(ind) * n_points evaluations go through one batched
        sweep instead of a Python loop.  Returns {param_index: (grid[n_points], values[n_points])}."""
        opt = np.asarray(param_opt, dtype=np.float64).reshape(-1)
        ind = list(range(opt.shape[0])) if ind is None else list(ind)
        grids, rows = [], []
        for a, j in enumerate(ind):
            if param_box is not None:
                lo, hi = float(np.asarray(param_box)[0, a]), float(np.asarray(param_box)[1, a])
            else:
                lo, hi = float(opt[j]) - float(delta), float(opt[j]) + float(delta)
            g = np.linspace(lo, hi, int(n_points))
            th = np.tile(opt, (int(n_points), 1))
            th[:, j] = g
            grids.append(g)
            rows.append(th)
        vals = self(np.concatenate(rows)) if rows else np.zeros(0)
        return {j: (grids[a], vals[a * int(n_points):(a + 1) * int(n_points)]) for a, j in enumerate(ind)}

    def profile_2d(self, covparam, param_indices=(0, 1), n=130, factor=10.0):
        """Criterion on the n x n grid of `plot_likelihood_sigma_rho`-style profiles
        (gpmp/modeldiagnosis/plotting.py:260-327): parameter i is swept geometrically by `factor` around its
        value in natural units (sigma = exp(theta_0 / 2) for index 0, rho = exp(-theta_i) otherwise).
        Returns (p1[n], p2[n], values[n, n]) with values[i, j] at (p1[j], p2[i]) like the reference's meshgrid;
        the n * n evaluations are one batched sweep."""
        cov0 = np.asarray(covparam, dtype=np.float64).reshape(-1)
        i1, i2 = param_indices

        def natural(i):
            return np.exp(cov0[i] / 2.0) if i == 0 else np.exp(-cov0[i])

        def to_log(i, mesh):
            return np.log(mesh ** 2) if i == 0 else np.log(1.0 / mesh)

        lf = np.log10(float(factor))
        p1 = np.logspace(np.log10(natural(i1)) - lf, np.log10(natural(i1)) + lf, int(n))
        p2 = np.logspace(np.log10(natural(i2)) - lf, np.log10(natural(i2)) + lf, int(n))
        m1, m2 = np.meshgrid(p1, p2)
        th = np.tile(cov0, (m1.size, 1))
        th[:, i1] = to_log(i1, m1).reshape(-1)
        th[:, i2] = to_log(i2, m2).reshape(-1)
        vals = np.nan_to_num(self(th)).reshape(m1.shape)
        return p1, p2, vals


class MiniBatchCriterion:
    """Selection criterion summed over mini-batches, all equal-size batches in ONE batched launch sequence.

    Mirrors `BatchDifferentiableSelectionCriterion` (gpmp/num/torch_backend.py:607-718): value = sum over
    batches of criterion(param, x_b, z_b) * len(batch), divided by the number of points for reduction="mean";
    `evaluate_pre_grad` caches the gradient that `gradient` returns.  `loader` is any iterable of (x_b, z_b)
    pairs (a torch DataLoader or a list); batches of the most common size are stacked and go through
    gpmp_criterion_batched_grad with one point set per entry, the remaining (ragged) batches through the scalar
    fused path.  The per-batch criterion is the model's REML (kind="reml", constant or zero mean shared by all
    batches) or zero-mean ML (kind="ml"); covariance = Matern-p with the package's parameter layout.
    """

    def __init__(self, model, loader, p, kind="reml", noise=False, reduction="mean", batches_per_eval=0):
        if reduction not in ("mean", "sum"):
            raise ValueError("reduction must be 'mean' or 'sum'")
        if batches_per_eval < 0:
            raise ValueError("batches_per_eval must be >= 0")
        if kind not in ("reml", "ml"):
            raise ValueError("kind must be 'reml' or 'ml'")
        if len(loader) == 0:
            raise ValueError("DataLoader is empty.")
        self.model, self.loader, self.p, self.kind, self.noise = model, loader, int(p), kind, bool(noise)
        self.reduction, self.bpe = reduction, int(batches_per_eval)
        self._batch_iter = iter(loader) if self.bpe > 0 else None
        self._gradient = None
        self._work, self._wkey = None, None

    def _batches(self):
        """The batches of one evaluation: the whole loader, or the next `batches_per_eval` ones (the loader is
        restarted when it runs out, so evaluations cycle through the data)."""
        if self.bpe == 0:
            return list(self.loader)
        picked = []
        while len(picked) < self.bpe:
            item = next(self._batch_iter, None)
            if item is None:
                self._batch_iter = iter(self.loader)
                continue
            picked.append(item)
        return picked

    def _basis(self, xb):
        if self.kind != "reml":
            return None
        P = _mean_values(self.model, xb)
        return (P.reshape(-1, 1) if P.dim() == 1 else P).contiguous()

    def _evaluate(self, param, want_grad):
        theta = ops.to_device(param).reshape(-1).detach()
        batches = [(ops.to_device(xb), ops.to_device(zb).reshape(-1)) for xb, zb in self._batches()]
        sizes = [b[0].shape[0] for b in batches]
        npts = sum(sizes)
        if npts == 0:
            raise ValueError("Loader is empty.")
        common = max(set(sizes), key=sizes.count)
        same = [b for b in batches if b[0].shape[0] == common]
        rest = [b for b in batches if b[0].shape[0] != common]
        total = torch.zeros((), dtype=torch.float64, device=theta.device)
        grad = torch.zeros_like(theta)
        if same:
            X = torch.stack([b[0] for b in same]).contiguous()
            Z = torch.stack([b[1] for b in same]).contiguous()
            P = self._basis(same[0][0])
            if P is not None and P.shape[1] > 0 and not all(torch.equal(self._basis(b[0]), P) for b in same[1:]):
                # basis depends on the points (e.g. linear mean): no shared P, take the scalar path for all
                rest, same = batches, []
            else:
                B, n, d = X.shape
                q = 0 if P is None else P.shape[1]
                key = (B, n, d, q)
                if self._wkey != key:
                    self._work, self._wkey = ops.criterion_batched_grad_workspace(n, q, d, B), key
                TH = theta.reshape(1, -1).repeat(B, 1)
                vals, grads, info = ops.criterion_batched_grad(TH, X, Z, P, self.p, self.noise, work=self._work)
                total = total + vals.sum() * common
                grad = grad + grads.sum(dim=0) * common
        for xb, zb in rest:
            tp = theta.clone().requires_grad_(want_grad)
            v = ops.fused_likelihood(tp, zb.contiguous(), xb.contiguous(), self._basis(xb), self.p, self.noise)
            if want_grad:
                (g,) = torch.autograd.grad(v, tp)
                grad = grad + g.to(grad.device) * xb.shape[0]
            total = total + v.detach().to(total.device) * xb.shape[0]
        if self.reduction == "mean":
            total, grad = total / npts, grad / npts
        return float(total), grad

    def evaluate(self, param):
        return self._evaluate(param, False)[0]

    def evaluate_no_grad(self, param):
        return self._evaluate(param, False)[0]

    def evaluate_pre_grad(self, param):
        value, grad = self._evaluate(param, True)
        self._gradient = grad.detach().cpu()
        return value

    def gradient(self, _param):
        if self._gradient is None:
            raise RuntimeError("Call `evaluate` first.")
        return self._gradient
