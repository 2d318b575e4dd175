"""Phase timestamps of the 128x128 diagonal-tile kernel (development aid)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gpmp_b200 import _abi
from oracle import gp_numpy as onp

L = _abi.lib()
fn = L.gpmp_debug_potf2
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_longlong, C.c_int] + [C.c_void_p] * 5
rng = np.random.default_rng(0)
x = rng.uniform(size=(128, 3))
K = onp.maternp_covariance(x, x, 2, np.array([0.0, 1.0, 1.0, 1.0])) + 1e-6 * np.eye(128)
for rep in range(3):
    A = torch.tensor(K, device="cuda")
    Tlo = torch.zeros(128, 128, dtype=torch.float64, device="cuda")
    Tup = torch.zeros_like(Tlo)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    clk = torch.zeros(32, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(A.data_ptr(), 128, 128, Tlo.data_ptr(), Tup.data_ptr(), info.data_ptr(), clk.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print("event ms", e0.elapsed_time(e1))
    c = clk.cpu().numpy()
names = {0: "start", 1: "loaded", 2: "p0 begin", 3: "p0 diag done", 4: "p0 rows solved", 5: "p1 begin", 6: "p1 diag",
         7: "p1 rows", 8: "p2 begin", 9: "p2 diag", 10: "p2 rows", 11: "p3 begin", 12: "p3 diag", 14: "factor done",
         15: "diag inverses", 16: "doubling 32", 17: "doubling 64", 18: "stored"}
prev = c[0]
for i in sorted(names):
    if c[i]:
        print(f"{names[i]:18s} +{c[i] - prev:8d} clk   t={c[i] - c[0]:8d}")
        prev = c[i]
err = np.abs(np.tril(A.cpu().numpy()) - np.linalg.cholesky(K)).max()
print("max err L", err, "info", info.item())
Tl = np.tril(Tlo.cpu().numpy())
print("max err T", np.abs(Tl - np.linalg.inv(np.linalg.cholesky(K))).max() / np.abs(Tl).max())
