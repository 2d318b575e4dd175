"""Two-rank NCCL test of the panel-partitioned factorisation and gradient (BASELINE config 5a path), run when at least
two GPUs are visible: `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`.  Each rank is one process
on one GPU (torch.multiprocessing spawn, rendezvous on 127.0.0.1); the partitioned value and gradient must equal the
single-GPU ones to rounding, on every rank, for a size with several column groups per rank and a ragged last group."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, d, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import gpmp_b200 as gp
    from oracle import cases

    x, z, _ = cases.data(n, d, 9)
    th = np.concatenate(([0.1], np.full(d, -np.log(0.6))))
    m = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                      lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, 2, cp, pairwise), covparam=th)
    xd, zd = gp.num.asarray(x), gp.num.asarray(z)
    tp = torch.tensor(th, requires_grad=True)
    v_loc = m.negative_log_restricted_likelihood(tp, xd, zd)
    (g_loc,) = torch.autograd.grad(v_loc, tp)
    res = []
    for _ in range(2):  # twice: buffers / streams / events are reused across calls
        v, g = gp.dist.reml_value_and_grad_distributed(m, th, xd, zd)
        res.append((abs(v - v_loc.item()) / abs(v_loc.item()),
                    float(np.max(np.abs(g - g_loc.numpy())) / np.max(np.abs(g_loc.numpy())))))
    xt = np.random.default_rng(3).uniform(size=(300, d))
    mean, var = gp.dist.predict_distributed(m, x, z, xt, fitted=gp.dist.fit_distributed(m, x, z))
    mean1, var1 = m.predict(x, z, xt)
    pe = float(max(np.max(np.abs(mean - mean1)), np.max(np.abs(var - var1))))
    out.put((rank, res, pe))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d", [(4500, 4), (8192 + 300, 6)])
def test_partitioned_value_and_gradient_two_ranks(n, d):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + (os.getpid() + n) % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, d, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res, pe in results:
        for ev, eg in res:
            print(f"[parity] partitioned n={n} rank {rank}: value {ev:.2e}, gradient {eg:.2e}, predict {pe:.2e}")
            assert ev <= 1e-11 and eg <= 1e-9
        assert pe <= 1e-10
