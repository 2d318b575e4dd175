"""BASELINE config 4, second half: a COMPLETE tempered SMC over the REML posterior of theta by the reference's own
sampler (gp.mcmc.sample_from_selection_criterion_smc -> gpmp/mcmc/smc.py, vendored unmodified under oracle/_ref)
with its particle loop (mcmc/param_posterior.py:752) bound to the batched sweep of this library.

    python scripts/smc_full.py [n_particles=8192] [n=512] [d=4] [mh_steps=5]
    python -m torch.distributed.run --nproc-per-node 8 ... scripts/smc_full.py      # theta rows sharded over ranks

The sampler itself (host control flow, unseeded generators: smc.py:129,535) runs on rank 0 only; before every sweep
rank 0 broadcasts the particle set, every rank evaluates its block of rows and the values are all-gathered
(gpmp_b200.dist.DrivenSweeps around a BatchedCriterion with a process group).
Rank 0 prints one JSON line: wall time, sweeps, sweeps/s, particle evaluations/s, posterior mean / std of theta.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cases, vendor_ref

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
group = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
d = int(sys.argv[3]) if len(sys.argv) > 3 else 4
mh = int(sys.argv[4]) if len(sys.argv) > 4 else 5

gp = vendor_ref.import_reference("torch")
import gpmp.num as gnp
import gpmp_b200.dropin as b200

b200.install(gp)
x, z, _ = cases.data(n, d, 77)
th0 = cases.theta(d, 77)
box = [list(th0 - 3.0), list(th0 + 3.0)]
model = gp.core.Model(lambda x_, param: gnp.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise),
                      None, None)
crit = b200.BatchableCriterion(model, x, z, 2, kind="reml", group=(dist.group.WORLD if world > 1 else None))
if world > 1:
    from gpmp_b200.dist import DrivenSweeps

    sweeps = DrivenSweeps(crit.crit, d + 1, dist.group.WORLD)  # crit.crit: the sharded BatchedCriterion
    if rank != 0:
        sweeps.serve()
        dist.destroy_process_group()
        sys.exit(0)
    crit.crit = sweeps
crit.batched(np.tile(th0, (N, 1)))  # warm-up: workspace, streams
crit.sweeps = crit.evaluations = 0
torch.cuda.synchronize()
t0 = time.perf_counter()
particles, smc = gp.mcmc.sample_from_selection_criterion_smc(
    selection_criterion=crit, init_box=box, sampling_box=box, n_particles=N, mh_steps=mh)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
P = gnp.to_np(particles)
# device time of one full sweep, for the split between sweeps and the sampler's own host work
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
crit.batched(P)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print(json.dumps({
        "workload": f"tempered SMC, {N} particles, REML at n={n}, d={d}, Matern p=2, {mh} MH moves per stage "
                    "(reference sampler, batched sweeps)",
        "n_gpus": world, "wall_s": wall, "sweeps": crit.sweeps, "evaluations": crit.evaluations,
        "sweeps_per_s": crit.sweeps / wall, "particle_evals_per_s": crit.evaluations / wall,
        "one_full_sweep_ms": e0.elapsed_time(e1),
        "reference_serial_s_per_sweep": "196 (BASELINE.md: 24 ms per evaluation x 8192, torch-CPU)",
        "posterior_mean": P.mean(axis=0).tolist(), "posterior_std": P.std(axis=0).tolist(),
        "theta_data_generating": th0.tolist()}))
b200.uninstall()
if world > 1:
    sweeps.stop()
    dist.destroy_process_group()
