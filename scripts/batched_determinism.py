import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases
n, d, p, N = 512, 4, 2, 8192
x, z, _ = cases.data(n, d, 77)
th0 = cases.theta(d, 77)
TH = th0 + np.random.default_rng(1).uniform(-2.0, 2.0, size=(N, d + 1))
m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
a = gp.batched.BatchedCriterion(m, x, z, p)(TH)
per = gp._abi.lib().gpmp_criterion_batched_bytes(n, 1, 2) - gp._abi.lib().gpmp_criterion_batched_bytes(n, 1, 1)
b = gp.batched.BatchedCriterion(m, x, z, p, max_bytes=1024 * per)(TH)
c = gp.batched.BatchedCriterion(m, x, z, p)(TH[:1024])
a2 = gp.batched.BatchedCriterion(m, x, z, p)(TH)
print("repeat equal", np.array_equal(a, a2), "chunk1024 equal", np.array_equal(a, b), "first1024 equal", np.array_equal(a[:1024], c))
diff = np.abs(a - b) / np.abs(a)
print("max rel", diff.max(), "n differing", int((a != b).sum()), "argmax", int(diff.argmax()))
idx = np.nonzero(a != b)[0][:10]
print(idx, a[idx], b[idx])
