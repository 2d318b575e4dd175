"""Sharding helpers: one process per GPU, torch.distributed for the plumbing.

The path shards only over independent units (particles / restarts / prediction points):
rows are block-partitioned, every rank runs the same kernels on its block, and the only collective is
the all-gather of the per-row results (NCCL on GPUs; gloo in the CPU tests of the partitioning logic).
"""
from __future__ import annotations

import torch
import torch.distributed as td


def world(group=None):
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


def block_bounds(N, rank, size):
    """[lo, hi) of `rank` in a balanced block partition of N rows (first N % size ranks get one more)."""
    base, rem = divmod(N, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bounds(N, group=None):
    rank, size = world(group)
    return block_bounds(N, rank, size)


def all_gather_rows(local, N, group=None):
    """Concatenate the ranks' blocks (block_bounds layout) of a tensor whose first dimension is sharded
    (values: 1-D; gradient rows: 2-D) into the full tensor with N rows."""
    rank, size = world(group)
    if size == 1:
        return local
    base, rem = divmod(N, size)
    width = base + (1 if rem else 0)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(size)]
    td.all_gather(parts, pad, group=group)
    chunks = []
    for r in range(size):
        lo, hi = block_bounds(N, r, size)
        chunks.append(parts[r][: hi - lo])
    return torch.cat(chunks)


def reml_value_distributed(model, covparam, xi, zi, group=None, want_grad=False):
    """Model.negative_log_restricted_likelihood (value only) with the Cholesky factorisation partitioned over
    the ranks of `group` (BASELINE config 5a).  Returns (value, FitState); every rank gets the same value and
    a complete fitted state (so prediction chunks can be sharded over ranks afterwards)."""
    from . import core, kernel, num, ops

    xi_, zi_, _ = core._ensure_shapes_and_type(xi=xi, zi=zi)
    P = model._basis(xi_) if model.meantype == "linear_predictor" else None
    if model.meantype == "parameterized":
        zi_ = zi_ - model.mean_values(xi_, num.asparam(model.meanparam)).reshape(-1)
    K, fused = model._same_set_cov(xi_, num.asparam(covparam))
    with torch.no_grad():
        if fused:
            spec = ops._spec_from_param(K.p, xi_.shape[1], ops.host_values(K.param))
            state, out = ops.lik_value_dist(spec, None, xi_, zi_, P, group, want_grad)
        else:
            Kd = kernel.materialize(K).detach().contiguous()
            state, out = ops.lik_value_dist(None, Kd, None, zi_, P, group, want_grad)
    return ops.read_small(out)[0], state


def reml_value_and_grad_distributed(model, covparam, xi, zi, group=None):
    """REML value and covariance-parameter gradient with both the factorisation (column groups) and the
    gradient's triangular inverse / K^-1 / contraction (row blocks) partitioned over the ranks of `group`.
    The covariance must be a plain gp.kernel.maternp_covariance (fused path).  Returns (value, grad ndarray),
    identical on every rank."""
    import math

    from . import ops

    value, state = reml_value_distributed(model, covparam, xi, zi, group, want_grad=True)
    nparam = len(ops.host_values(covparam))
    if not math.isfinite(value):
        import numpy as np

        return value, np.zeros(nparam)
    with torch.no_grad():
        g, _alpha = ops.lik_grad_dist(state, group)
    vals = ops._expand_iso_grad(ops.read_small(g), nparam, state.d, 1)
    import numpy as np

    return value, np.asarray(vals)


def fit_distributed(model, xi, zi, group=None):
    """Model.fit with the factorisation partitioned over the ranks of `group`: every rank returns a `Fitted`
    handle holding the complete factor (model.covparam / meanparam are used, as in Model.predict)."""
    _, state = reml_value_distributed(model, model.covparam, xi, zi, group)
    return model.fit(xi, zi, state=state)


def predict_distributed(model, xi, zi, xt, group=None, convert_out=True, fitted=None):
    """Model.predict with the test points block-partitioned over the ranks of `group` (independent columns,
    SURVEY.md 8e).  `fitted`: a handle from fit_distributed / Model.fit (otherwise every rank fits locally);
    each rank predicts its block of xt and the (mean, variance) vectors are all-gathered."""
    from . import num, ops

    rank, size = world(group)
    xt = ops.to_device(xt)
    m = xt.shape[0]
    lo, hi = block_bounds(m, rank, size)
    if fitted is None:
        fitted = model.fit(xi, zi)
    mean, var = fitted.predict(xt[lo:hi], convert_out=False)
    mean = all_gather_rows(mean, m, group)
    var = all_gather_rows(var, m, group)
    if convert_out:
        return num.to_np(mean), num.to_np(var)
    return mean, var


class DrivenSweeps:
    """One rank drives, the others serve: for samplers whose host control flow cannot run in lock step on every
    rank (GPmp's SMC / MH draw from unseeded generators, gpmp/mcmc/smc.py:129,535), rank 0 runs the sampler and
    calls `sweeps(rows)` for every particle set; the other ranks sit in `sweeps.serve()`.  Before each sweep rank 0
    broadcasts the row count and the rows, then EVERY rank calls `fn(rows)` -- a collective function that evaluates
    its own block of rows and all-gathers the results (BatchedCriterion with a process group) -- so the particle
    loop of mcmc/param_posterior.py:752 is sharded although only one rank holds the sampler's state.

        sweeps = DrivenSweeps(crit, dim, group)       # crit: BatchedCriterion(..., group=group)
        if rank != 0: sweeps.serve()                  # returns when rank 0 calls sweeps.stop()
        else:
            with sweeps: ...sampler calling sweeps(thetas)...      # stop() on exit, also when the sampler raises
    """

    def __init__(self, fn, dim, group=None, device=None):
        self.fn, self.dim, self.group = fn, int(dim), group
        self.rank, self.size = world(group)
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self.src = td.get_global_rank(group, 0) if (group is not None and self.size > 1) else 0
        self._stopped = False

    def __call__(self, rows):
        import numpy as np

        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.dim)
        if self.size > 1:
            if self.rank != 0:
                raise RuntimeError("DrivenSweeps: only rank 0 drives; the other ranks call serve()")
            td.broadcast(torch.tensor([rows.shape[0]], dtype=torch.int64, device=self.device), src=self.src,
                         group=self.group)
            td.broadcast(torch.as_tensor(rows, device=self.device), src=self.src, group=self.group)
        return self.fn(rows)

    def serve(self):
        """Ranks != 0: evaluate sweeps until rank 0 sends the stop word; returns the number of sweeps served."""
        served = 0
        while self.size > 1:
            hdr = torch.zeros(1, dtype=torch.int64, device=self.device)
            td.broadcast(hdr, src=self.src, group=self.group)
            cnt = int(hdr.item())
            if cnt < 0:
                break
            buf = torch.empty((cnt, self.dim), dtype=torch.float64, device=self.device)
            td.broadcast(buf, src=self.src, group=self.group)
            self.fn(buf.cpu().numpy())
            served += 1
        return served

    def stop(self):
        if self.size > 1 and self.rank == 0 and not self._stopped:
            self._stopped = True
            td.broadcast(torch.tensor([-1], dtype=torch.int64, device=self.device), src=self.src, group=self.group)

    # `with DrivenSweeps(...) as sweeps:` on rank 0 releases the serving ranks even if the sampler raises
    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.stop()
        return False
