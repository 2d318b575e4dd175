"""One REML value+gradient at a given n through the public API (short command for ncu captures)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x, z, th0 = cases.headline(n=n)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
model = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise))
for _ in range(reps):
    tp = torch.tensor(th0, requires_grad=True)
    v = model.negative_log_restricted_likelihood(tp, xd, zd)
    (g,) = torch.autograd.grad(v, tp)
torch.cuda.synchronize()
print(v.item(), g.numpy())
