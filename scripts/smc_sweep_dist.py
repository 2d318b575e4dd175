"""Particle-batched REML sweep (BASELINE config 4: 8192 particles x n=512, d=4) sharded over the ranks of a
torchrun launch; rank 0 prints sweeps/s and particle-evals/s (max-over-ranks device time)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import gpmp_b200 as gp
from oracle import cases

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, p, N = 512, 4, 2, 8192
x, z, _ = cases.data(n, d, 77)
th0 = cases.theta(d, 77)
TH = th0 + np.random.default_rng(1).uniform(-2.0, 2.0, size=(N, d + 1))
m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise))
crit = gp.batched.BatchedCriterion(m, x, z, p, kind="reml")
for _ in range(2):
    vals = crit(TH)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    vals = crit(TH)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"n_gpus": world, "sweep_ms": ms, "particle_evals_per_s": N / ms * 1e3,
                      "tflops": N * n**3 / 3 / ms / 1e9, "finite": float(np.isfinite(vals).mean()),
                      "checksum": float(np.sum(vals[np.isfinite(vals)]))}))
if world > 1:
    dist.destroy_process_group()
