"""Development: per-launch timeline (class, stream, start, duration) of one REML value at n."""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import ops, _abi
from oracle import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, z, th0 = cases.headline(n=n)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
P = gp.num.ones((n, 1))
spec = _abi.make_spec(2, x.shape[1], th0[0], th0[1:])
for _ in range(3):
    ops.lik_value(spec, None, xd, zd, P, False)
torch.cuda.synchronize()
want_grad = len(sys.argv) > 4 and sys.argv[4] == "grad"
if want_grad:
    st, _ = ops.lik_value(spec, None, xd, zd, P, True)
    ops.lik_grad(st, False, False)
    torch.cuda.synchronize()
_abi.prof_enable(True)
if want_grad:
    st, _ = ops.lik_value(spec, None, xd, zd, P, True)
    torch.cuda.synchronize()
    _abi.prof_enable(False); [_abi.prof_read(c) for c in range(6)]; _abi.prof_enable(True)
    ops.lik_grad(st, False, False)
else:
    # "withgrad": the forward pass as the value+gradient path runs it (gradient-sized workspace: the leading block
    # of T = L^-1 is computed under the tail of the factorisation)
    ops.lik_value(spec, None, xd, zd, P, len(sys.argv) > 4 and sys.argv[4] == "withgrad")
torch.cuda.synchronize()
lib = _abi.lib()
buf = (C.c_double * (4 * 4000))()
lib.gpmp_debug_timeline.restype = C.c_int
cnt = lib.gpmp_debug_timeline(buf, 4000)
_abi.prof_enable(False)
rows = np.array(buf[: 4 * cnt]).reshape(cnt, 4)
names = ["matern", "gemm", "potf2", "contract", "small", "batched"]
order = np.argsort(rows[:, 2], kind="stable")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else cnt
print("launches", cnt, "span ms", rows[:, 2].max() + rows[np.argmax(rows[:, 2]), 3])
for i in order[lo:hi]:
    c, sid, st, du = rows[i]
    print(f"{st*1e3:9.1f} us  +{du*1e3:7.1f}  stream {int(sid)}  {names[int(c)]}")
