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


class BatchedCriterion:
    def __init__(self, model, xi, zi, p, kind="reml", noise=False, group=None, max_bytes=None):
        if kind not in ("reml", "ml"):
            raise ValueError("kind must be 'reml' or 'ml'")
        self.x = ops.to_device(xi)
        z = ops.to_device(zi).reshape(-1)
        self.p, self.noise, self.kind = int(p), bool(noise), kind
        self.P = None
        if kind == "reml":
            P = ops.to_device(model.mean(self.x, model.meanparam))
            self.P = (P.reshape(-1, 1) if P.dim() == 1 else P).contiguous()
        elif model.meantype == "parameterized":
            z = z - ops.to_device(model.mean(self.x, model.meanparam)).reshape(-1)
        self.z = z.contiguous()
        self.group = group
        self.max_bytes = max_bytes
        self._work = None  # device workspace, allocated at the first sweep and reused

    def values_device(self, thetas):
        """thetas: (N, 1+noise+d) array-like -> (values, info) device tensors for the local shard."""
        th = ops.to_device(thetas)
        n = self.x.shape[0]
        q = 0 if self.P is None else self.P.shape[1]
        if self._work is None or self._rows < th.shape[0]:
            self._work = ops.criterion_batched_workspace(n, q, th.shape[0], self.max_bytes)
            self._rows = th.shape[0]
        return ops.criterion_batched(th, self.x, self.z, self.P, self.p, self.noise, self.max_bytes, self._work)

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
