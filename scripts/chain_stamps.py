"""Development: phase clocks of the fused chain-step kernel (narrow CTA and wide CTA 0) of the last stamped launches
of one REML value at n (run with GPMP_DEV_STAMPS=1)."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import ops, _abi
from oracle import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x, z, th0 = cases.headline(n=n)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
P = gp.num.ones((n, 1))
spec = _abi.make_spec(2, x.shape[1], th0[0], th0[1:])
for _ in range(3):
    ops.lik_value(spec, None, xd, zd, P, False)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
lib = _abi.lib()
lib.gpmp_debug_chain_stamps.restype = C.c_int
assert lib.gpmp_debug_chain_stamps(buf) == 0
v = np.array(buf[:64])
names_n = ["start", "staged", "solved", "published", "syrk done", "packed", "factored", "stored"]
print("narrow CTA (clk from start, delta):")
for i in range(1, 8):
    print(f"  {names_n[i]:10s} {v[i]-v[0]:8d}  +{v[i]-v[i-1]:7d}")
print("wide CTA 0:")
w = v[16:32]
for i in range(1, 13):
    if w[i] > w[0]:
        print(f"  stamp {i:2d} {w[i]-w[0]:8d}")
