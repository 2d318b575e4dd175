"""REML value at n (no gradient): device time per evaluation and the host time spent enqueueing it."""
import sys, time
import torch
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
    state, out = ops.lik_value(spec, None, xd, zd, P, False)
torch.cuda.synchronize()
host, dev = [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    state, out = ops.lik_value(spec, None, xd, zd, P, False)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
print("n", n, "host enqueue ms", [round(h, 2) for h in host], "device ms", [round(d, 2) for d in dev])
