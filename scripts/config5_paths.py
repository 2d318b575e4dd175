"""BASELINE config 5b at a bounded size: n=32768, d=10; chunked predict + conditioning of `paths` sample paths
at m test points on the ranks of a torchrun launch (factorisation partitioned over ranks, xt rows sharded)."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
paths = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
d = 10
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x, z, _ = cases.data(n, d, 9)
th = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
model = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise), covparam=th)
lo, hi = gp.dist.block_bounds(m, rank, world)
ml = hi - lo
gen = torch.Generator(device="cuda").manual_seed(100 + rank)
xt = torch.rand(ml, d, dtype=torch.float64, device="cuda", generator=gen)
# synthetic unconditional paths (throughput only): rows 0..n-1 = observation sites, n.. = this rank's test sites
ztsim = torch.randn(n + ml, paths, dtype=torch.float64, device="cuda", generator=gen)
xd, zd = gp.num.asarray(x), gp.num.asarray(z)


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


sync(); t0 = time.perf_counter()
fitted = gp.dist.fit_distributed(model, xd, zd)
sync(); t_fit = time.perf_counter() - t0
sync(); t0 = time.perf_counter()
mean, var = fitted.predict(xt, convert_out=False)
sync(); t_pred = time.perf_counter() - t0
sync(); t0 = time.perf_counter()
cond = fitted.conditional_sample_paths_chunked(ztsim, np.arange(n), xt, n + np.arange(ml), convert_out=False)
sync(); t_cond = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"n": n, "m_total": m, "paths": paths, "world": world, "fit_s": t_fit, "predict_s": t_pred,
                      "predict_points_per_s": m / t_pred, "predict_tflops": float(n) * n * m / t_pred / 1e12,
                      "cond_s": t_cond, "cond_points_per_s": m / t_cond,
                      "cond_tflops": (float(n) * n * m + 2.0 * n * m * paths) / t_cond / 1e12,
                      "var_min": float(var.min()), "finite": bool(torch.isfinite(cond).all())}))
if world > 1:
    dist.destroy_process_group()
