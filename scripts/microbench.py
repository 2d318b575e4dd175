"""Device microbenchmarks used while developing (not part of the judged bench): FP64 GEMM peak through
torch/cuBLAS vs the library's DMMA GEMM, and the stage timings of one REML value+gradient evaluation."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import _abi, ops
from oracle import cases


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


out = {}
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, dtype=torch.float64, device="cuda")
tmin, tmed = ev_time(lambda: torch.matmul(A, B.t()), reps=10)
out["cublas_dgemm_8192_tflops_best"] = 2 * n**3 / tmin / 1e9
out["cublas_dgemm_8192_tflops_median"] = 2 * n**3 / tmed / 1e9
Cc = torch.empty(n, n, dtype=torch.float64, device="cuda")
tmin, tmed = ev_time(lambda: ops.gemm_nt(A, B, C_out=Cc), reps=10)
out["dmma_gemm_8192_tflops_best"] = 2 * n**3 / tmin / 1e9
out["dmma_gemm_8192_tflops_median"] = 2 * n**3 / tmed / 1e9
err = (Cc - torch.matmul(A, B.t())).abs().max().item()
out["dmma_vs_cublas_maxabs"] = err
for k in (128, 256, 512):
    Ak, Bk = A[:, :k].contiguous(), B[:, :k].contiguous()
    tmin, _ = ev_time(lambda: ops.gemm_nt(Ak, Bk, C_out=Cc, alpha=-1.0, beta=1.0, lower=True), reps=10)
    out[f"dmma_syrk_8192_k{k}_tflops"] = n * n * k / tmin / 1e9
del A, B, Cc
torch.cuda.empty_cache()

# stage timings of the headline evaluation
x, z, th0 = cases.headline()
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
model = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise))


def value_only():
    with torch.no_grad():
        return model.negative_log_restricted_likelihood(th0, xd, zd)


def value_grad():
    tp = torch.tensor(th0, requires_grad=True)
    v = model.negative_log_restricted_likelihood(tp, xd, zd)
    (g,) = torch.autograd.grad(v, tp)
    return v, g


for name, fn in (("value", value_only), ("value_grad", value_grad)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    out[f"reml_{name}_ms_median"] = float(np.median(ts))
    out[f"reml_{name}_ms_min"] = float(min(ts))

_abi.prof_enable(True)
value_grad()
torch.cuda.synchronize()
names = ["matern", "gemm", "potf2", "contract", "small", "batched"]
for c, nm in enumerate(names):
    ms, cnt, work = _abi.prof_read(c)
    out[f"prof_{nm}"] = {"ms": ms, "launches": cnt, "work": work}
_abi.prof_enable(False)

# torch/cuSOLVER potrf for context
K = gp.kernel.maternp_covariance(xd, None, 2, th0)
tmin, _ = ev_time(lambda: torch.linalg.cholesky(K), reps=3, warm=1)
out["cusolver_potrf_8192_ms"] = tmin
tmin, _ = ev_time(lambda: ops.potrf(K, check_pd=False), reps=3, warm=1)
out["gpmp_potrf_8192_ms_with_copy"] = tmin
print(json.dumps(out, indent=1))
