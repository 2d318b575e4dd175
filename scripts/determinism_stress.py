"""Repeat the pipelined value+gradient at several sizes and check bit-identical results run to run
(a missing stream dependency would show up as run-to-run differences) and agreement with a size-independent
identity (d/d log sigma2 closed form)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

for n, d in [(8192, 8), (5000, 5), (4700, 3), (6273, 6), (12289, 4)]:
    x, z, th = cases.headline(n=n, d=d)
    m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise))
    xd, zd = gp.num.asarray(x), gp.num.asarray(z)
    vals, grads = [], []
    for rep in range(8):
        tp = torch.tensor(th, requires_grad=True)
        v = m.negative_log_restricted_likelihood(tp, xd, zd)
        (g,) = torch.autograd.grad(v, tp)
        vals.append(v.item()); grads.append(g.numpy().copy())
    same_v = all(v == vals[0] for v in vals)
    same_g = all(np.array_equal(g, grads[0]) for g in grads)
    quad = m.norm_k_sqrd(xd, zd, th).item()
    closed = 0.5 * ((n - 1) - quad)
    print(f"n={n} d={d} identical values {same_v} gradients {same_g}  value {vals[0]:.12g}  "
          f"|g0 - closed form| {abs(grads[0][0] - closed):.2e}")
