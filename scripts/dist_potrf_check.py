"""Panel-partitioned REML value over the ranks of a torchrun launch vs the single-GPU value (same inputs)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, z, _ = cases.data(n, d, 9)
th = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise), covparam=th)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)
with torch.no_grad():
    v1 = m.negative_log_restricted_likelihood(th, xd, zd).item()
vd, state = gp.dist.reml_value_distributed(m, th, xd, zd)
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    vd, state = gp.dist.reml_value_distributed(m, th, xd, zd)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
t_local = []
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        m.negative_log_restricted_likelihood(th, xd, zd)
    torch.cuda.synchronize()
    t_local.append(time.perf_counter() - t0)
# gradient: partitioned vs local autograd
tp = torch.tensor(th, requires_grad=True)
v_loc = m.negative_log_restricted_likelihood(tp, xd, zd)
(g_loc,) = torch.autograd.grad(v_loc, tp)
vg, gd = gp.dist.reml_value_and_grad_distributed(m, th, xd, zd)
tg = []
for _ in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    vg, gd = gp.dist.reml_value_and_grad_distributed(m, th, xd, zd)
    torch.cuda.synchronize()
    tg.append(time.perf_counter() - t0)
tl = []
for _ in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tp = torch.tensor(th, requires_grad=True)
    v_loc = m.negative_log_restricted_likelihood(tp, xd, zd)
    (g_loc,) = torch.autograd.grad(v_loc, tp)
    torch.cuda.synchronize()
    tl.append(time.perf_counter() - t0)
grad_rel = float(np.max(np.abs(gd - g_loc.numpy())) / np.max(np.abs(g_loc.numpy())))
if rank == 0:
    print(json.dumps({"grad_rel_dist_vs_local": grad_rel, "t_value_grad_dist_s": min(tg), "t_value_grad_local_s": min(tl),
                      "tflops_vg_dist": float(n) ** 3 / min(tg) / 1e12, "tflops_vg_local": float(n) ** 3 / min(tl) / 1e12}))
if rank == 0:
    print(json.dumps({"n": n, "world": world, "value_local": v1, "value_dist": vd, "rel": abs(vd - v1) / abs(v1),
                      "t_dist_s": min(ts), "t_local_s": min(t_local),
                      "tflops_dist": n**3 / 3 / min(ts) / 1e12, "tflops_local": n**3 / 3 / min(t_local) / 1e12}))
if world > 1:
    dist.destroy_process_group()
