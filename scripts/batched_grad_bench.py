"""Timing of the batched value+gradient criterion (config-4 shape) and of a mini-batch epoch."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

n, d, p = 512, 4, 2
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, z, _ = cases.data(n, d, 77)
th0 = cases.theta(d, 77)
TH = th0 + np.random.default_rng(1).uniform(-1.0, 1.0, size=(N, d + 1))
m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise), None, th0)
crit = gp.batched.BatchedCriterion(m, x, z, p)
out = {"n": n, "d": d, "N": N}
for _ in range(2):
    crit.value_and_grad(TH, convert_out=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    crit.value_and_grad(TH, convert_out=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
out["batched_value_grad_ms"] = ms
out["tflops"] = N * float(n) ** 3 / (ms * 1e-3) / 1e12
# scalar loop for comparison (8 particles)
f = lambda t: m.negative_log_restricted_likelihood(t, x, z)
gp.num.value_and_grad(f, TH[0]); torch.cuda.synchronize()
t0 = time.time()
for i in range(8):
    gp.num.value_and_grad(f, TH[i])
torch.cuda.synchronize()
out["scalar_value_grad_ms_each"] = (time.time() - t0) / 8 * 1e3
# mini-batch epoch: 65536 points in batches of 512
rng = np.random.default_rng(5)
X = rng.uniform(size=(65536, d)); Z = np.sin(3 * X.sum(1)) + 0.1 * rng.standard_normal(65536)
Xd, Zd = gp.num.asarray(X), gp.num.asarray(Z)
loader = [(Xd[i:i + 512], Zd[i:i + 512]) for i in range(0, 65536, 512)]
mb = gp.batched.MiniBatchCriterion(m, loader, p)
mb.evaluate_pre_grad(th0); torch.cuda.synchronize()
t0 = time.time()
for _ in range(3):
    mb.evaluate_pre_grad(th0)
torch.cuda.synchronize()
out["minibatch_epoch_128x512_ms"] = (time.time() - t0) / 3 * 1e3
print(json.dumps(out))
