"""Host-side logic that needs no GPU: block partitioning of particles over ranks (world_size-2 gloo run),
the bench harness contract (reference arm, input generator), the build entry point."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpmp_b200 import dist as gdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_bounds_cover_exactly():
    for N in (0, 1, 7, 8, 8191, 8192):
        for size in (1, 2, 3, 8):
            got = []
            for r in range(size):
                lo, hi = gdist.block_bounds(N, r, size)
                assert 0 <= lo <= hi <= N and hi - lo in (N // size, N // size + 1)
                got.extend(range(lo, hi))
            assert got == list(range(N))


def _worker(rank, world, port, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gdist.shard_bounds(N)
    local = torch.arange(lo, hi, dtype=torch.float64) ** 2  # stands for the rank's criterion values
    full = gdist.all_gather_rows(local, N)
    # gradient rows (2-D, first dimension sharded) travel the same way
    rows = torch.stack([torch.arange(lo, hi, dtype=torch.float64), -torch.arange(lo, hi, dtype=torch.float64)], dim=1)
    full_rows = gdist.all_gather_rows(rows, N)
    assert full_rows.shape == (N, 2) and torch.equal(full_rows[:, 0], torch.arange(N, dtype=torch.float64))
    assert torch.equal(full_rows[:, 1], -torch.arange(N, dtype=torch.float64))
    q.put((rank, lo, hi, full.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [5, 8, 13])
def test_particle_sharding_world2_gloo(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + N) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [float(i * i) for i in range(N)]
    for rank, lo, hi, full in res:
        assert full == want
    assert sorted((lo, hi) for _, lo, hi, _ in res) == [gdist.block_bounds(N, 0, 2), gdist.block_bounds(N, 1, 2)]


def test_bench_inputs_are_the_seeded_headline_case():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import cases

    x, z, th = bench.headline_inputs(64, 8)
    xr, zr, thr = cases.headline(64, 8)
    assert np.array_equal(x, xr) and np.array_equal(z, zr) and np.array_equal(th, thr)
    assert bench.METRIC.startswith("REML logL+grad evals/s")


def test_bench_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-n", "1024"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["metric"] == "REML logL+grad evals/s (n=8192,d=8,fp64)"


def test_build_entry_point():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    assert os.path.exists(os.path.join(ROOT, "gpmp_b200", "libgpmp_b200.so"))
