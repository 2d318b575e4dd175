"""DMMA GEMM throughput at several K (development aid; the tile shapes it once compared are listed in gemm.cu)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from gpmp_b200 import ops


def ev(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
out = {}
out["gemm_8192"] = 2 * n**3 / ev(lambda: ops.gemm_nt(A, B, C_out=C)) / 1e9
for k in (128, 256, 512, 1024):
    Ak = A[:, :k].contiguous()
    out[f"syrk_k{k}"] = n * n * k / ev(lambda: ops.gemm_nt(Ak, Ak, C_out=C, alpha=-1.0, beta=1.0, lower=True)) / 1e9
out["lauum_like"] = (n**3 / 3) / ev(lambda: ops.gemm_nt(A, A, C_out=C, tri=1, lower=True)) / 1e9
print(json.dumps(out))
