"""Two DMMA GEMM launches for ncu: a K=4096 product and a K=512 lower-triangular SYRK update (n=8192)."""
import sys

import torch

sys.path.insert(0, ".")
from gpmp_b200 import ops

n = 8192
A = torch.randn(n, 4096, dtype=torch.float64, device="cuda")
C = torch.zeros(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    ops.gemm_nt(A[:4096], A[4096:], C_out=C[:4096, :4096])
    ops.gemm_nt(A[:, :512], A[:, :512], C_out=C, alpha=-1.0, beta=1.0, lower=True)
torch.cuda.synchronize()
print("ok")
