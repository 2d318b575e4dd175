"""One REML value (no gradient) at n through the public API: the forward chain only (for ncu launch lists)."""
import sys

import torch

sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x, z, th0 = cases.headline(n=n)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
model = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise))
for _ in range(2):
    with torch.no_grad():
        v = model.negative_log_restricted_likelihood(th0, xd, zd)
torch.cuda.synchronize()
print(v.item())
