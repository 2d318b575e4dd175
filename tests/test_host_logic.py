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


# ---- host-side selection driver (gpmp_b200/selection.py; mirrors kernel/parameter_selection.py:128-276) ----------
def test_autoselect_parameters_driver_on_a_quadratic():
    """The SciPy loop around a criterion: bounds, history, best-seen tracking and the info fields the samplers
    read -- exercised on a closed-form criterion (no device needed)."""
    from gpmp_b200 import selection

    target = np.array([0.5, -1.5, 2.0])

    def crit(p):
        return float(np.sum((np.asarray(p) - target) ** 2) + 3.0)

    def grad(p):
        return 2.0 * (np.asarray(p) - target)

    p0 = np.zeros(3)
    for method in ("SLSQP", "L-BFGS-B"):
        best, info = selection.autoselect_parameters(p0, crit, grad, info=True, method=method)
        assert np.allclose(best, target, atol=1e-4) and abs(info.fun - 3.0) <= 1e-6
        assert len(info["history_params"]) == len(info["history_criterion"]) >= 2
        assert np.array_equal(info["initial_params"], p0) and np.allclose(info["final_params"], best)
        assert info["bounds"] == [(-10.0, 10.0)] * 3 and info["total_time"] >= 0.0
    # explicit bounds are honoured; automatic bounds are clipped to +-500
    best, _ = selection.autoselect_parameters(p0, crit, grad, bounds=[(-1, 0.2), (-1, 1), (0, 1)])
    assert np.allclose(best, [0.2, -1.0, 1.0], atol=1e-6)
    _, info = selection.autoselect_parameters(np.array([495.0]), lambda p: float((p[0] - 490.0) ** 2),
                                              lambda p: 2.0 * (np.asarray(p) - 490.0), info=True)
    assert info["bounds"] == [(485.0, 500.0)]
    with pytest.raises(ValueError):
        selection.autoselect_parameters(p0, crit, grad, method="Nelder-Mead")


def test_autoselect_parameters_maps_linear_algebra_failures_to_inf():
    """A criterion that raises a linear-algebra error on part of the domain counts as +inf there
    (parameter_selection.py:222-231); any other exception propagates."""
    from gpmp_b200 import selection

    def crit(p):
        if p[0] > 1.0:
            raise torch.linalg.LinAlgError("not positive definite")
        return float((p[0] - 0.9) ** 2)

    best, info = selection.autoselect_parameters(np.array([0.0]), crit, lambda p: 2.0 * (np.asarray(p) - 0.9),
                                                 info=True)
    assert abs(best[0] - 0.9) <= 1e-3 and np.isfinite(info["history_criterion"]).any()

    def broken(p):
        raise KeyError("unrelated")

    with pytest.raises(KeyError):
        selection.autoselect_parameters(np.array([0.0]), broken, lambda p: np.zeros(1))


# ---- mini-batch criterion: host logic (grouping by batch size, reductions, cycling) without a device ----------
def test_minibatch_criterion_host_logic_with_oracle_ops(monkeypatch):
    """MiniBatchCriterion stacks equal-size batches into one batched call and sends the ragged batch through the
    scalar path; here both device entry points are replaced by the CPU oracle, so the grouping, the weighting
    by batch size, both reductions and `batches_per_eval` are checked on CPU against the reference's own
    BatchDifferentiableSelectionCriterion (golden vectors)."""
    from gpmp_b200 import batched, ops
    from oracle import cases, gp_torch as otorch
    from oracle.make_golden_minibatch import MINIBATCH_CASES

    calls = {"batched": 0, "entries": 0, "scalar": 0}

    def fake_batched(theta, x, z, P, p, noise=False, max_bytes=None, work=None):
        calls["batched"] += 1
        calls["entries"] += theta.shape[0]
        vals, grads = [], []
        for b in range(theta.shape[0]):
            v, g = otorch.reml_value_and_grad(x[b], z[b], P, p, theta[b], noise)
            vals.append(v)
            grads.append(g)
        return (torch.tensor(vals, dtype=torch.float64), torch.tensor(np.array(grads), dtype=torch.float64),
                torch.zeros(theta.shape[0], dtype=torch.int32))

    def fake_scalar(param, z, x, P, p, noise=False):
        calls["scalar"] += 1
        if P is None:
            return otorch.nll_zero_mean(x, z, p, param, noise)
        return otorch.reml(x, z, P, p, param, noise)

    monkeypatch.setattr(ops, "to_device", lambda a, requires_contiguous=True: torch.as_tensor(a, dtype=torch.float64))
    monkeypatch.setattr(ops, "criterion_batched_grad", fake_batched)
    monkeypatch.setattr(ops, "criterion_batched_grad_workspace", lambda n, q, d, N, max_bytes=None: None)
    monkeypatch.setattr(ops, "fused_likelihood", fake_scalar)

    class _Model:
        def __init__(self, kind):
            self.mean, self.meanparam = cases.mean_fn(kind, torch), None
            self.meantype = cases.meantype_of(kind)

    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_minibatch.npz"))
    for name, n, bs, d, p, kind, seed in MINIBATCH_CASES:
        x, z, th = g[name + "/x"], g[name + "/z"], g[name + "/theta"]
        loader = [(x[i:i + bs], z[i:i + bs]) for i in range(0, n, bs)]
        nfull, ragged = n // bs, int(n % bs != 0)
        for red in ("mean", "sum"):
            calls.update(batched=0, entries=0, scalar=0)
            c = batched.MiniBatchCriterion(_Model(kind), loader, p, kind="ml" if kind == "zero" else "reml",
                                           reduction=red)
            v = c.evaluate_pre_grad(th)
            assert calls == {"batched": 1, "entries": nfull, "scalar": ragged}
            assert abs(v - float(g[f"{name}/{red}/value"])) <= 1e-9 * abs(v)
            gr = np.asarray(c.gradient(th))
            ref = g[f"{name}/{red}/grad"]
            assert np.linalg.norm(gr - ref) <= 1e-8 * np.linalg.norm(ref)
        # two batches per evaluation, cycling through the loader
        c = batched.MiniBatchCriterion(_Model(kind), loader, p, kind="ml" if kind == "zero" else "reml",
                                       batches_per_eval=2)
        calls.update(batched=0, entries=0, scalar=0)
        for _ in range(len(loader)):
            c.evaluate(th)
        assert calls["entries"] + calls["scalar"] == 2 * len(loader)
    with pytest.raises(ValueError):
        batched.MiniBatchCriterion(_Model("const"), [], 2)
    with pytest.raises(ValueError):
        batched.MiniBatchCriterion(_Model("const"), loader, 2, reduction="median")
