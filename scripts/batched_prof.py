import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import _abi
from oracle import cases
n, d, p, N = 512, 4, 2, 8192
x, z, _ = cases.data(n, d, 77)
th0 = cases.theta(d, 77)
TH = th0 + np.random.default_rng(1).uniform(-2.0, 2.0, size=(N, d + 1))
m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
crit = gp.batched.BatchedCriterion(m, x, z, p)
thd = gp.num.asarray(TH)
for _ in range(2):
    crit.values_device(thd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); crit.values_device(thd); e1.record(); torch.cuda.synchronize()
out = {"sweep_ms": e0.elapsed_time(e1)}
_abi.prof_enable(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
crit.values_device(thd)
torch.cuda.synchronize()
for c, nm in enumerate(["matern", "gemm", "potf2", "contract", "small", "batched"]):
    ms, cnt, work = _abi.prof_read(c)
    out[nm] = {"ms": round(ms, 3), "launches": cnt}
_abi.prof_enable(False)
print(json.dumps(out))
