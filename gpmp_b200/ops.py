"""torch custom ops (torch.autograd.Function) over the C-ABI of libgpmp_b200.so.

Every op here launches hand-written sm_100a kernels through ``_abi``; there is no PyTorch-math or CPU
fallback.  torch is used for device memory (tensors as buffers), the current stream and autograd glue.

Reference interfaces replaced (paths relative to the GPmp 0.9.37 source tree):
  scaled_distance / scaled_distance_elementwise   gpmp/num/torch_backend.py:810-829
  maternp_kernel / maternp_covariance             gpmp/kernel/matern.py:32-141
  cholesky / cholesky_solve / cholesky_inv / ...  gpmp/num/torch_backend.py:111,841-890
  the three likelihoods and their gradients       gpmp/core/likelihood.py:18-129 (+ autograd)
  kriging predictor                               gpmp/core/kriging.py:35-116,170-199
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _abi
from ._abi import check, lib, ptr, stream_ptr

F64 = torch.float64


# --------------------------------------------------------------------------------------------------
# device plumbing
# --------------------------------------------------------------------------------------------------
def device():
    _abi.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def to_device(a, requires_contiguous=True):
    """numpy / list / CPU tensor / CUDA tensor -> float64 CUDA tensor (no copy when already there)."""
    if a is None:
        return None
    dev = device()
    if torch.is_tensor(a):
        t = a
        if t.dtype != F64:
            t = t.to(F64)
        if t.device != dev:
            t = t.to(dev, non_blocking=True)
    else:
        t = torch.as_tensor(a, dtype=F64)
        if t.device.type == "cpu":
            # stage through pinned memory so the copy is asynchronous w.r.t. the host
            t = t.contiguous().pin_memory().to(dev, non_blocking=True)
        else:
            t = t.to(dev)
    if requires_contiguous and not t.is_contiguous():
        t = t.contiguous()
    return t


def host_values(t):
    """Small parameter tensor -> list of python floats (host)."""
    if torch.is_tensor(t):
        return t.detach().reshape(-1).to("cpu", F64).tolist()
    return [float(v) for v in torch.as_tensor(t, dtype=F64).reshape(-1).tolist()]


def _empty(shape, dtype=F64):
    return torch.empty(shape, dtype=dtype, device=device())


def _workspace(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device())


_pinned = {}


def _pinned_buf(n, key="d"):
    buf = _pinned.get((key, n))
    if buf is None:
        buf = torch.empty(n, dtype=F64).pin_memory()
        _pinned[(key, n)] = buf
    return buf


def read_small(dev_tensor, n=None):
    """Device float64 vector -> python list through a pinned buffer (one D2H copy + stream sync)."""
    n = dev_tensor.numel() if n is None else n
    buf = _pinned_buf(n)
    buf.copy_(dev_tensor.reshape(-1)[:n], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return buf.tolist()


def _ld(t):
    return t.stride(0) if t.dim() == 2 else t.shape[0]


def _spec_from_param(p, d, values, noise=False):
    """values = [log s2, (log tau2,) loginvrho...] (python floats)."""
    if noise:
        return _abi.make_spec(p, d, values[0], values[2:], noise=True, log_tau2=values[1])
    return _abi.make_spec(p, d, values[0], values[1:])


def _param_grad_like(param, grad_list):
    g = torch.tensor(grad_list, dtype=F64)
    if param.device.type != "cpu":
        g = g.to(param.device)
    return g.reshape(param.shape).to(param.dtype)


def _expand_iso_grad(grad, nparam, d, head):
    """Fold the d per-dimension gradients back when loginvrho was a scalar (isotropic)."""
    if nparam == head + d:
        return grad
    return grad[:head] + [sum(grad[head:])]


# --------------------------------------------------------------------------------------------------
# L0: distances
# --------------------------------------------------------------------------------------------------
class _ScaledDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loginvrho, x, y):
        lir = host_values(loginvrho)
        n, d = x.shape
        if len(lir) == 1 and d > 1:
            lir = lir * d
        same = y is None
        m = n if same else y.shape[0]
        D = _empty((n, m))
        arr = (C.c_double * d)(*lir)
        if n and m:
            check(lib().gpmp_scaled_distance(arr, d, ptr(x), n, ptr(None if same else y), m, ptr(D), _ld(D),
                                             stream_ptr()), "gpmp_scaled_distance")
        ctx.meta = (lir, d, n, m, same, loginvrho.numel() if torch.is_tensor(loginvrho) else len(lir))
        ctx.save_for_backward(x, y if not same else x, loginvrho if torch.is_tensor(loginvrho) else x)
        return D

    @staticmethod
    def backward(ctx, G):
        lir, d, n, m, same, nparam = ctx.meta
        x, y, lref = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        G = G.contiguous()
        nbytes = lib().gpmp_contract_workspace_bytes(n, m, d)
        ws = _workspace(nbytes)
        g = _empty((1 + d,))
        arr = (C.c_double * d)(*lir)
        check(lib().gpmp_scaled_distance_backward(arr, d, ptr(x), n, ptr(None if same else y), m, ptr(G), _ld(G),
                                                  ptr(g), ptr(ws), ws.numel(), stream_ptr()),
              "gpmp_scaled_distance_backward")
        vals = read_small(g)[1:]
        if nparam == 1 and d > 1:
            vals = [sum(vals)]
        return _param_grad_like(lref, vals), None, None


def scaled_distance(loginvrho, x, y):
    """D_ik = || exp(loginvrho) * (x_i - y_k) ||; `y is x or y is None` gives the exact-zero diagonal."""
    xd = to_device(x)
    same = y is None or y is x
    yd = None if same else to_device(y)
    if not torch.is_tensor(loginvrho):
        loginvrho = torch.as_tensor(loginvrho, dtype=F64)
    return _ScaledDistance.apply(loginvrho, xd, yd)


def scaled_distance_elementwise(loginvrho, x, y):
    xd = to_device(x)
    n, d = xd.shape
    same = y is None or y is x
    out = _empty((n,))
    lir = host_values(loginvrho)
    if len(lir) == 1 and d > 1:
        lir = lir * d
    arr = (C.c_double * d)(*lir)
    yd = None if same else to_device(y)
    if n:
        check(lib().gpmp_scaled_distance_elementwise(arr, d, ptr(xd), ptr(yd), n, ptr(out), stream_ptr()),
              "gpmp_scaled_distance_elementwise")
    return out


# --------------------------------------------------------------------------------------------------
# L1: Matern kernel / covariance
# --------------------------------------------------------------------------------------------------
class _MaternKernel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, p):
        hc = h.contiguous()
        k = torch.empty_like(hc)
        need = ctx.needs_input_grad[0]
        dk = torch.empty_like(hc) if need else None
        if hc.numel():
            check(lib().gpmp_maternp_kernel(int(p), ptr(hc), ptr(k), ptr(dk), hc.numel(), stream_ptr()),
                  "gpmp_maternp_kernel")
        if need:
            ctx.save_for_backward(dk)
        return k

    @staticmethod
    def backward(ctx, g):
        (dk,) = ctx.saved_tensors
        return g * dk, None


def maternp_kernel(p, h):
    """k_p(h) = exp(-c h) q_p(2 c h), c = 2 sqrt(p + 1/2) (kernel/matern.py:32-64), elementwise on the device."""
    return _MaternKernel.apply(to_device(h), int(p))


class _MaternCov(torch.autograd.Function):
    """Full covariance matrix sigma2 k_p(D) (+ nugget I on the same set) with a dK-regenerating backward."""

    @staticmethod
    def forward(ctx, param, x, y, p):
        vals = host_values(param)
        n, d = x.shape
        spec = _spec_from_param(p, d, vals)
        same = y is None
        m = n if same else y.shape[0]
        K = _empty((n, m))
        if n and m:
            check(lib().gpmp_matern_cov(C.byref(spec), ptr(x), n, ptr(y), m, ptr(K), _ld(K), _abi.COV_FULL,
                                        stream_ptr()), "gpmp_matern_cov")
        ctx.meta = (spec, n, m, d, same, len(vals))
        ctx.save_for_backward(x, x if same else y, param)
        return K

    @staticmethod
    def backward(ctx, G):
        spec, n, m, d, same, nparam = ctx.meta
        x, y, param = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        G = G.contiguous()
        ws = _workspace(lib().gpmp_contract_workspace_bytes(n, m, d))
        g = _empty((1 + d,))
        check(lib().gpmp_matern_cov_backward(C.byref(spec), ptr(x), n, ptr(None if same else y), m, ptr(G), _ld(G),
                                             ptr(g), ptr(ws), ws.numel(), stream_ptr()), "gpmp_matern_cov_backward")
        vals = _expand_iso_grad(read_small(g), nparam, d, 1)
        return _param_grad_like(param, vals), None, None, None


def matern_cov(x, y, p, param):
    xd = to_device(x)
    same = y is None or y is x
    yd = None if same else to_device(y)
    if not torch.is_tensor(param):
        param = torch.as_tensor(param, dtype=F64)
    return _MaternCov.apply(param, xd, yd, int(p))


def matern_cov_pairwise(x, y, p, param):
    xd = to_device(x)
    n, d = xd.shape
    same = y is None or y is x
    spec = _spec_from_param(p, d, host_values(param))
    out = _empty((n,))
    yd = None if same else to_device(y)
    if n:
        check(lib().gpmp_matern_cov_pairwise(C.byref(spec), ptr(xd), ptr(yd), n, ptr(out), stream_ptr()),
              "gpmp_matern_cov_pairwise")
    return out


# --------------------------------------------------------------------------------------------------
# L0: Cholesky family
# --------------------------------------------------------------------------------------------------
class Factor:
    """A Cholesky factorisation living on the device: the both-ways factor matrix and the workspace of
    diagonal-block inverses that gpmp_potri / gpmp_trsm_rows start from."""

    def __init__(self, A, n, work, info):
        self.A, self.n, self.work, self.info = A, n, work, info

    def lower(self):
        return torch.tril(self.A[: self.n, : self.n])


def _round_ld(n):
    return (n + 15) // 16 * 16


def potrf(K, extra_rows=None, check_pd=True):
    """Factor the symmetric matrix K (its lower triangle is read).  extra_rows (r x n) are whitened along."""
    K = to_device(K)
    n = K.shape[0]
    r = 0 if extra_rows is None else extra_rows.shape[0]
    lda = _round_ld(n)
    A = _empty((n + r, lda))
    A[:n, :n].copy_(K)
    if r:
        A[n:, :n].copy_(extra_rows)
    nbytes = lib().gpmp_potrf_workspace_bytes(n, n + r)
    work = _workspace(nbytes)
    info = torch.zeros(1, dtype=torch.int32, device=device())
    if n:
        check(lib().gpmp_potrf(ptr(A), n, n + r, lda, ptr(work), work.numel(), ptr(info), stream_ptr()), "gpmp_potrf")
    if check_pd:
        bad = int(info.item())
        if bad:
            raise torch.linalg.LinAlgError(
                f"gpmp_potrf: the input is not positive-definite (the leading minor of order {bad} is not "
                "positive-definite)")
    return Factor(A, n, work, info)


def trsm_rows(fac, Bt, trans):
    """Rows of Bt (m x n) solved in place: trans=0 -> L^-1 b, trans=1 -> L^-T b."""
    m = Bt.shape[0]
    scratch = _workspace(max(m, 1) * 512 * 8)
    check(lib().gpmp_trsm_rows(ptr(fac.A), fac.n, _ld(fac.A), ptr(fac.work), ptr(Bt), m, _ld(Bt), int(trans),
                               ptr(scratch), stream_ptr()), "gpmp_trsm_rows")
    return Bt


def potri(fac):
    """(Kinv_lower, Tlo, Tup) from a factor."""
    n = fac.n
    ld = _round_ld(n)
    Tlo, Tup, Kinv = _empty((n, ld)), _empty((n, ld)), _empty((n, ld))
    check(lib().gpmp_potri(ptr(fac.A), n, _ld(fac.A), ptr(fac.work), ptr(Tlo), ptr(Tup), ptr(Kinv), ld,
                           stream_ptr()), "gpmp_potri")
    return Kinv, Tlo, Tup


def gemm_nt(A, B, C_out=None, alpha=1.0, beta=0.0, tri=0, lower=False):
    """C = alpha A B^T + beta C on the DMMA pipe; A (M x K), B (N x K) row-major with even leading dims."""
    M, K = A.shape
    N = B.shape[0]
    if C_out is None:
        C_out = _empty((M, _round_ld(N)))[:, :N]
    check(lib().gpmp_gemm_nt(ptr(A), _ld(A), ptr(B), _ld(B), ptr(C_out), _ld(C_out), M, N, K, float(alpha),
                             float(beta), int(tri), int(bool(lower)), stream_ptr()), "gpmp_gemm_nt")
    return C_out


def padded(t):
    """Copy of a 2-D tensor whose leading dimension is a multiple of 16 elements (GEMM operand)."""
    r, c = t.shape
    buf = _empty((r, _round_ld(c)))
    out = buf[:, :c]
    out.copy_(t)
    return out


def transpose(t):
    """Device transpose through the library's tiled kernel; result has a padded leading dimension."""
    t = t if t.stride(1) == 1 else t.contiguous()
    r, c = t.shape
    buf = _empty((c, _round_ld(r)))
    out = buf[:, :r]
    if r and c:
        check(lib().gpmp_transpose(ptr(t), _ld(t), ptr(out), _ld(out), r, c, stream_ptr()), "gpmp_transpose")
    return out


# --------------------------------------------------------------------------------------------------
# L2: likelihoods
# --------------------------------------------------------------------------------------------------
class FitState:
    """Device state left by gpmp_lik_value: L (both ways), whitened rows, R~; consumed by the gradient
    and by predict."""

    def __init__(self, work, n, q, d, spec, x, with_grad):
        self.work, self.n, self.q, self.d, self.spec, self.x, self.with_grad = work, n, q, d, spec, x, with_grad


def lik_value(spec, K, x, z, P, want_grad):
    """Enqueue the value pipeline; returns (FitState, out_dev[8])."""
    n = z.shape[0]
    q = 0 if P is None else P.shape[1]
    d = x.shape[1] if x is not None else 1
    if q > _abi.MAX_Q:
        raise _abi.GpmpError(f"mean basis has {q} columns; at most {_abi.MAX_Q} are supported")
    nbytes = lib().gpmp_lik_workspace_bytes(n, q, d, 1 if want_grad else 0)
    work = _workspace(nbytes)
    out = _empty((8,))
    info = torch.empty(1, dtype=torch.int32, device=device())
    check(lib().gpmp_lik_value(C.byref(spec) if spec is not None else None, ptr(K), _ld(K) if K is not None else 0,
                               ptr(x), n, ptr(z), ptr(P), q, ptr(work), work.numel(), ptr(out), ptr(info),
                               stream_ptr()), "gpmp_lik_value")
    return FitState(work, n, q, d, spec, x, want_grad), out


def lik_value_dist(spec, K, x, z, P, group=None, want_grad=False):
    """lik_value with the factorisation partitioned over the ranks of `group` (one process per GPU): column
    groups are owned round-robin, the owner factors a group, the panel is broadcast with NCCL, every rank
    updates the groups it owns.  Every rank ends with the same FitState a local lik_value would leave."""
    import torch.distributed as td

    from . import dist as gdist

    rank, size = gdist.world(group)
    n = z.shape[0]
    q = 0 if P is None else P.shape[1]
    d = x.shape[1] if x is not None else 1
    work = _workspace(lib().gpmp_lik_workspace_bytes(n, q, d, 1 if want_grad else 0))
    out = _empty((8,))
    info = torch.empty(1, dtype=torch.int32, device=device())
    specp = C.byref(spec) if spec is not None else None
    check(lib().gpmp_lik_dist_prepare(specp, ptr(K), _ld(K) if K is not None else 0, ptr(x), n, ptr(z), ptr(P), q,
                                      ptr(work), work.numel(), ptr(info), stream_ptr()), "gpmp_lik_dist_prepare")
    NB = lib().gpmp_lik_dist_block(n)
    nrows = n + q + 1
    ngroups = (n + NB - 1) // NB
    # Three panel buffers in rotation and two streams per rank.  The CHAIN stream (high priority) carries what the
    # other ranks are waiting for -- the owner's update of its next group, that group's factorisation and the start
    # of its broadcast; the caller's stream carries the rank's bulk updates.  A latency-bound group factorisation
    # therefore no longer holds back the owner's own trailing updates, and a panel buffer is only rewritten once
    # the bulk updates that read it (three panels earlier) have finished.
    NBUF = 3
    bufs = [_empty((nrows * NB,)) for _ in range(NBUF)]
    wb = work.numel()
    main = torch.cuda.current_stream()
    chain = _chain_stream() if size > 1 else main
    done_with = [None] * ngroups   # event: the bulk updates reading panel g are enqueued and finished
    head_ready = [None] * ngroups  # event: group g has received every bulk update issued so far (through panel g - 2)

    def root(g):
        owner = g % size
        return td.get_global_rank(group, owner) if group is not None else owner

    def panel(g):
        return bufs[g % NBUF][: (nrows - g * NB) * NB]

    def update(g, g2):
        check(lib().gpmp_lik_dist_update(n, q, ptr(work), wb, g * NB, ptr(bufs[g % NBUF]), g2 * NB,
                                         min(n, (g2 + 1) * NB), stream_ptr()), "gpmp_lik_dist_update")

    def factor(g):
        check(lib().gpmp_lik_dist_group(n, q, ptr(work), wb, g * NB, ptr(bufs[g % NBUF]), ptr(info), stream_ptr()),
              "gpmp_lik_dist_group")

    def start_broadcast(g):
        """On the chain stream: the buffer of panel g must be free (its previous tenant, panel g - NBUF, fully
        consumed by this rank's bulk updates), then the collective is enqueued behind the stream's work."""
        if size == 1:
            return None
        if g >= NBUF and done_with[g - NBUF] is not None:
            chain.wait_event(done_with[g - NBUF])
        with torch.cuda.stream(chain):
            return td.broadcast(panel(g), src=root(g), group=group, async_op=True)

    chain.wait_stream(main)  # the prepared work matrix
    if rank == 0 % size:
        with torch.cuda.stream(chain):
            factor(0)
    pending = start_broadcast(0)
    for g in range(ngroups):
        if pending is not None:
            pending.wait()  # panel g has arrived: the caller's stream waits for it ...
            with torch.cuda.stream(chain):
                pending.wait()  # ... and so does the chain
        elif size == 1:
            pass
        if size > 1 and rank == g % size:
            main.wait_stream(chain)  # the owner's own factorisation produced the panel on the chain stream
        if size > 1 and rank != g % size:
            check(lib().gpmp_lik_dist_store(n, q, ptr(work), wb, g * NB, ptr(bufs[g % NBUF]), stream_ptr()),
                  "gpmp_lik_dist_store")
        nxt = g + 1
        if nxt < ngroups:
            if rank == nxt % size:
                # the chain: bring the next group up to date, factor it, start its broadcast
                with torch.cuda.stream(chain):
                    if head_ready[nxt] is not None:
                        chain.wait_event(head_ready[nxt])
                    if g >= NBUF - 1 and done_with[nxt - NBUF] is not None:
                        chain.wait_event(done_with[nxt - NBUF])  # factor(nxt) rewrites that buffer
                    update(g, nxt)
                    factor(nxt)
            pending = start_broadcast(nxt)
        # bulk updates of the groups this rank owns; the group the chain needs next goes first and is flagged
        mine = [g2 for g2 in range(g + 2, ngroups) if g2 % size == rank]
        for g2 in mine:
            update(g, g2)
            if g2 == g + 2:
                head_ready[g2] = torch.cuda.Event()
                head_ready[g2].record(main)
        done_with[g] = torch.cuda.Event()
        done_with[g].record(main)
    main.wait_stream(chain)
    if size > 1:
        # a non-positive pivot is seen by the owner of its group only
        td.all_reduce(info, op=td.ReduceOp.MAX, group=group)
    check(lib().gpmp_lik_dist_finish(n, q, ptr(work), wb, ptr(out), ptr(info), stream_ptr()),
          "gpmp_lik_dist_finish")
    return FitState(work, n, q, d, spec, x, want_grad), out


_chain_streams = {}


def _chain_stream():
    """High-priority stream of the current device for the critical path of the partitioned factorisation."""
    dev = torch.cuda.current_device()
    st = _chain_streams.get(dev)
    if st is None:
        st = torch.cuda.Stream(device=dev, priority=-1)
        _chain_streams[dev] = st
    return st


def _ws_matrix(state, which):
    """(n x ld) float64 view of a matrix inside a likelihood workspace (0 A, 1 T^T, 2 K^-1, 3 U, 4 T)."""
    n, q = state.n, state.q
    off = lib().gpmp_lik_ws_offset(n, q, which)
    ld = lib().gpmp_lik_ws_ld(n)
    rows = (q + 1) if which == 3 else n
    return state.work[off: off + rows * ld * 8].view(torch.float64).view(rows, ld)


def lik_grad_dist(state, group=None):
    """Covariance-parameter gradient (and dvalue/dz) of a state left by lik_value_dist(want_grad=True), with
    the 2n^3/3 of T = L^-1 and K^-1 split over the ranks by row blocks (owned round-robin like the column
    groups of the factorisation): T^T rows -> broadcast -> K^-1 rows, U columns -> all-reduce -> contraction of
    the owned rows -> all-reduce of the 1+d partial sums."""
    import torch.distributed as td

    from . import dist as gdist

    rank, size = gdist.world(group)
    n, q, d, spec = state.n, state.q, state.d, state.spec
    if spec is None:
        raise _abi.GpmpError("the partitioned gradient needs a Matern covariance spec (fused path)")
    wb = state.work.numel()
    NB = lib().gpmp_lik_dist_block(n)
    blocks = [(r0, min(NB, n - r0)) for r0 in range(0, n, NB)]
    mine = [b for i, b in enumerate(blocks) if i % size == rank]
    Tup, U = _ws_matrix(state, 1), _ws_matrix(state, 3)
    for r0, rows in mine:
        check(lib().gpmp_lik_dist_tup_rows(n, q, ptr(state.work), wb, r0, rows, stream_ptr()),
              "gpmp_lik_dist_tup_rows")
    if size > 1:
        pend = []
        for i, (r0, rows) in enumerate(blocks):
            src = td.get_global_rank(group, i % size) if group is not None else i % size
            pend.append(td.broadcast(Tup[r0:r0 + rows], src=src, group=group, async_op=True))
        for h in pend:
            h.wait()
    U.zero_()
    for r0, rows in mine:
        check(lib().gpmp_lik_dist_kinv_rows(n, q, ptr(state.work), wb, r0, rows, stream_ptr()),
              "gpmp_lik_dist_kinv_rows")
        check(lib().gpmp_lik_dist_u_cols(n, q, ptr(state.work), wb, r0, rows, stream_ptr()), "gpmp_lik_dist_u_cols")
    if size > 1:
        td.all_reduce(U, op=td.ReduceOp.SUM, group=group)
    ng = 1 + spec.noise + d
    total = torch.zeros(ng, dtype=F64, device=device())
    g = _empty((ng,))
    for r0, rows in mine:
        check(lib().gpmp_lik_dist_contract_rows(C.byref(spec), ptr(state.x), n, q, ptr(state.work), wb, r0, rows,
                                                ptr(g), stream_ptr()), "gpmp_lik_dist_contract_rows")
        total += g
    if size > 1:
        td.all_reduce(total, op=td.ReduceOp.SUM, group=group)
    return total, U[q, :n].clone()


def lik_grad(state, want_dz, want_dK):
    n, q, d = state.n, state.q, state.d
    spec = state.spec
    ng = (1 + spec.noise + d) if spec is not None else 0
    g = _empty((max(ng, 1),))
    dz = _empty((n,)) if want_dz else None
    dK = _empty((n, n)) if want_dK else None
    check(lib().gpmp_lik_grad(C.byref(spec) if spec is not None else None, ptr(state.x), n, q, ptr(state.work),
                              state.work.numel(), ptr(g), ptr(dz), ptr(dK), n if want_dK else 0, stream_ptr()),
          "gpmp_lik_grad")
    return g, dz, dK


def lik_loo(state, z):
    """(zloo, sigma2loo, eloo) device vectors from a fitted state (workspace sized for the gradient)."""
    n = state.n
    zloo, s2, e = _empty((n,)), _empty((n,)), _empty((n,))
    check(lib().gpmp_lik_loo(n, state.q, ptr(state.work), state.work.numel(), ptr(z), ptr(zloo), ptr(s2), ptr(e),
                             stream_ptr()), "gpmp_lik_loo")
    return zloo, s2, e


def lik_trsm_rows(state, Bt, trans):
    """Row solves (trans=0: b -> L^-1 b, trans=1: b -> L^-T b) against the factor held in a fitted state."""
    m = Bt.shape[0]
    scratch = _workspace(max(m, 1) * 512 * 8)
    check(lib().gpmp_lik_trsm_rows(state.n, state.q, ptr(state.work), state.work.numel(), ptr(Bt), m, _ld(Bt),
                                   int(trans), ptr(scratch), stream_ptr()), "gpmp_lik_trsm_rows")
    return Bt


def _scalar_like(param, value):
    t = torch.tensor(value, dtype=F64)
    if torch.is_tensor(param) and param.device.type != "cpu":
        t = t.to(param.device)
    return t


class _FusedLikelihood(torch.autograd.Function):
    """criterion(param, z) for a Matern-p covariance on x with mean basis P (REML, q >= 1) or without
    (zero-mean ML, q == 0); K is built, factored and contracted entirely on the device."""

    @staticmethod
    def forward(ctx, param, z, x, P, p, noise):
        vals = host_values(param)
        n, d = x.shape
        spec = _spec_from_param(p, d, vals, noise=noise)
        want_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        state, out = lik_value(spec, None, x, z, P, want_grad)
        host = read_small(out)
        ctx.state = state
        ctx.nparam = len(vals)
        ctx.head = 2 if noise else 1
        ctx.value = host[0]
        ctx.save_for_backward(param, z)
        return _scalar_like(param, host[0])

    @staticmethod
    def backward(ctx, gout):
        param, z = ctx.saved_tensors
        state = ctx.state
        if not math.isfinite(ctx.value):
            gp = torch.zeros_like(param) if ctx.needs_input_grad[0] else None
            gz = torch.zeros_like(z) if ctx.needs_input_grad[1] else None
            return gp, gz, None, None, None, None
        g, dz, _ = lik_grad(state, want_dz=ctx.needs_input_grad[1], want_dK=False)
        gp = gz = None
        scale = float(gout)
        if ctx.needs_input_grad[0]:
            vals = _expand_iso_grad(read_small(g), ctx.nparam, state.d, ctx.head)
            gp = _param_grad_like(param, [scale * v for v in vals])
        if ctx.needs_input_grad[1]:
            gz = dz * scale
        return gp, gz, None, None, None, None


class _LikelihoodFromK(torch.autograd.Function):
    """criterion(K, z) for a user-composed covariance matrix K (lower triangle read); backward emits the
    dense dvalue/dK = 0.5 (Pi - alpha alpha^T) so autograd can continue through the user's ops."""

    @staticmethod
    def forward(ctx, K, z, P):
        want_grad = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        Kc = K if K.stride(1) == 1 else K.contiguous()
        state, out = lik_value(None, Kc, None, z, P, want_grad)
        host = read_small(out)
        ctx.state = state
        ctx.value = host[0]
        ctx.kshape = K.shape
        return torch.tensor(host[0], dtype=F64, device=K.device)

    @staticmethod
    def backward(ctx, gout):
        state = ctx.state
        if not math.isfinite(ctx.value):
            n = state.n
            gK = torch.zeros(ctx.kshape, dtype=F64, device=gout.device) if ctx.needs_input_grad[0] else None
            gz = torch.zeros(n, dtype=F64, device=gout.device) if ctx.needs_input_grad[1] else None
            return gK, gz, None
        _, dz, dK = lik_grad(state, want_dz=ctx.needs_input_grad[1], want_dK=ctx.needs_input_grad[0])
        gK = dK * gout if dK is not None else None
        gz = dz * gout if dz is not None else None
        return gK, gz, None


def fused_likelihood(param, z, x, P, p, noise=False):
    if not torch.is_tensor(param):
        param = torch.as_tensor(param, dtype=F64)
    return _FusedLikelihood.apply(param, z, x, P, int(p), bool(noise))


def likelihood_from_K(K, z, P):
    return _LikelihoodFromK.apply(K, z, P)


# --------------------------------------------------------------------------------------------------
# L2: prediction
# --------------------------------------------------------------------------------------------------
def predict_chunk(state, xt, Pt, ktt, Vt, want_lambda, return_dots=False):
    """One chunk of test points against a fitted state; returns (mean, var) device vectors; Vt is
    overwritten: want_lambda = 0 rows v_t = L^-1 k(xi, x_t) (the whitened cross-covariances), 1 rows lambda_t^T,
    2 rows w_t = v_t - Q~ e_t (the kriging weights in whitened coordinates: lambda_t = L^-T w_t).
    return_dots: also return the per-point record the row pass leaves in the scratch, (m, q + 2):
    e_t = Q~^T v_t - R~^-T p_t (q entries), v_t . r, |v_t|^2  (posterior covariances need e_t)."""
    m = Vt.shape[0]
    n, q = state.n, state.q
    sbytes = lib().gpmp_predict_scratch_bytes(n, q, m)
    scratch = _workspace(sbytes)
    mean, var = _empty((m,)), _empty((m,))
    spec = state.spec
    check(lib().gpmp_predict_chunk(C.byref(spec) if spec is not None else None, ptr(state.x), n, q, ptr(state.work),
                                   state.work.numel(), ptr(xt), m, ptr(Pt), ptr(ktt), ptr(Vt), _ld(Vt),
                                   ptr(scratch), scratch.numel(), ptr(mean), ptr(var), int(want_lambda),
                                   stream_ptr()), "gpmp_predict_chunk")
    if return_dots:
        NB = lib().gpmp_lik_dist_block(n)
        off = (m * NB * 8 + 255) // 256 * 256
        dots = scratch[off: off + m * (q + 2) * 8].view(torch.float64).view(m, q + 2)
        return mean, var, dots
    return mean, var


# --------------------------------------------------------------------------------------------------
# batched criterion
# --------------------------------------------------------------------------------------------------
def criterion_batched_workspace(n, q, N, max_bytes=None):
    """Workspace tensor for up to N particles in flight (capped at max_bytes, default 16 GiB)."""
    if max_bytes is None:
        max_bytes = 16 << 30
    per1 = lib().gpmp_criterion_batched_bytes(n, q, 1)
    per2 = lib().gpmp_criterion_batched_bytes(n, q, 2) - per1
    nb = max(1, min(max(N, 1), (max_bytes - per1) // max(per2, 1) + 1))
    return _workspace(lib().gpmp_criterion_batched_bytes(n, q, int(nb)))


def _batched_operands(theta, x, z, P, noise):
    """Device, dtype, contiguity and shape checks shared by the batched entry points (the C side strides theta by
    1 + noise + d and trusts every pointer).  An isotropic parameter row [log s2, (log tau2,) log(1/rho)] is
    expanded to d length-scale columns; the flag tells the caller to fold the gradient back."""
    theta = to_device(theta)
    x = to_device(x)
    z = to_device(z)
    P = to_device(P)
    if theta.dim() != 2:
        raise _abi.GpmpError(f"theta must be a 2-D (N, 1+noise+d) array, got shape {tuple(theta.shape)}")
    if x.dim() not in (2, 3):
        raise _abi.GpmpError(f"x must be (n, d) or (N, n, d), got shape {tuple(x.shape)}")
    d = x.shape[-1]
    if d > _abi.MAX_DIM:
        raise _abi.GpmpError(f"input dimension {d} exceeds {_abi.MAX_DIM}")
    head = 1 + int(bool(noise))
    iso = False
    if theta.shape[1] == head + 1 and d > 1:
        theta = torch.cat((theta[:, :head], theta[:, head:].expand(-1, d)), dim=1).contiguous()
        iso = True
    if theta.shape[1] != head + d:
        raise _abi.GpmpError(f"theta has {theta.shape[1]} columns; expected 1 + noise + d = {head + d} "
                             f"(or {head + 1} for an isotropic kernel)")
    if P is not None:
        if P.dim() != 2 or P.shape[0] != x.shape[-2]:
            raise _abi.GpmpError(f"mean basis must be (n, q) with n = {x.shape[-2]}, got {tuple(P.shape)}")
        if P.shape[1] > _abi.MAX_Q:
            raise _abi.GpmpError(f"mean basis has {P.shape[1]} columns; at most {_abi.MAX_Q} are supported")
    return theta, x, z, P, iso


def criterion_batched(theta, x, z, P, p, noise=False, max_bytes=None, work=None):
    """N criterion values at the rows of theta (N x (1+noise+d)) on fixed (x, z, P); (values[N], info[N]) device
    tensors out.  `work` (optional) is a reusable workspace from criterion_batched_workspace; it decides how many
    particles are in flight per chunk."""
    theta, x, z, P, _ = _batched_operands(theta, x, z, P, noise)
    if x.dim() != 2 or z.dim() != 1 or z.shape[0] != x.shape[0]:
        raise _abi.GpmpError("criterion_batched takes one shared point set x (n, d) and observations z (n,)")
    N = theta.shape[0]
    n, d = x.shape
    q = 0 if P is None else P.shape[1]
    spec = _abi.make_spec(p, d, 0.0, [0.0] * d, noise=noise)
    values = _empty((N,))
    info = torch.empty(max(N, 1), dtype=torch.int32, device=device())
    if N == 0:
        return values, info[:0]
    if work is None:
        work = criterion_batched_workspace(n, q, N, max_bytes)
    check(lib().gpmp_criterion_batched(C.byref(spec), ptr(theta), N, ptr(x), n, ptr(z), ptr(P), q, ptr(work),
                                       work.numel(), ptr(values), ptr(info), stream_ptr()),
          "gpmp_criterion_batched")
    return values, info[:N]


def criterion_batched_grad_workspace(n, q, d, N, max_bytes=None):
    """Workspace for up to N entries of the batched value+gradient pipeline in flight (default cap 16 GiB)."""
    if max_bytes is None:
        max_bytes = 16 << 30
    per1 = lib().gpmp_criterion_batched_grad_bytes(n, q, d, 1)
    per2 = lib().gpmp_criterion_batched_grad_bytes(n, q, d, 2) - per1
    nb = max(1, min(max(N, 1), (max_bytes - per1) // max(per2, 1) + 1))
    return _workspace(lib().gpmp_criterion_batched_grad_bytes(n, q, d, int(nb)))


def criterion_batched_grad(theta, x, z, P, p, noise=False, max_bytes=None, work=None):
    """N criterion values and gradients in one call.  theta: (N, 1+noise+d); x: (n, d) shared points or
    (N, n, d) one point set per entry; z: (n,) or (N, n) likewise; P: (n, q) shared basis or None.
    Returns device tensors (values[N], grads[N, 1+noise+d], info[N]); gradient rows of entries that are not
    positive definite are zeroed (torch_backend.py:528-529)."""
    theta, x, z, P, iso = _batched_operands(theta, x, z, P, noise)
    N = theta.shape[0]
    per_entry_x = x.dim() == 3
    n, d = x.shape[-2], x.shape[-1]
    if per_entry_x and x.shape[0] != N:
        raise _abi.GpmpError(f"x has {x.shape[0]} point sets for {N} parameter rows")
    per_entry_z = z.dim() == 2
    if per_entry_z and z.shape[0] != N:
        raise _abi.GpmpError(f"z has {z.shape[0]} rows for {N} parameter rows")
    if z.shape[-1] != n:
        raise _abi.GpmpError(f"z has {z.shape[-1]} observations for {n} points")
    q = 0 if P is None else P.shape[1]
    spec = _abi.make_spec(p, d, 0.0, [0.0] * d, noise=noise)
    head = 1 + int(bool(noise))
    width = head + d
    values, grads = _empty((N,)), _empty((N, width))
    info = torch.empty(max(N, 1), dtype=torch.int32, device=device())
    if N == 0:
        return values, grads[:, : head + 1] if iso else grads, info[:0]
    if work is None:
        work = criterion_batched_grad_workspace(n, q, d, N, max_bytes)
    check(lib().gpmp_criterion_batched_grad(C.byref(spec), ptr(theta), N, ptr(x),
                                            n * d if per_entry_x else 0, n, ptr(z), n if per_entry_z else 0,
                                            ptr(P), q, ptr(work), work.numel(), ptr(values), ptr(grads), ptr(info),
                                            stream_ptr()), "gpmp_criterion_batched_grad")
    info = info[:N]
    grads = torch.where((info != 0).reshape(-1, 1), torch.zeros_like(grads), grads)
    if iso:  # one length-scale parameter: its gradient is the sum over the d expanded columns
        grads = torch.cat((grads[:, :head], grads[:, head:].sum(dim=1, keepdim=True)), dim=1)
    return values, grads, info
