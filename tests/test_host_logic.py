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


def _driven_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def sharded(rows):  # what BatchedCriterion.__call__ does, with a CPU stand-in for the device sweep
        N = rows.shape[0]
        lo, hi = gdist.shard_bounds(N)
        local = torch.as_tensor((rows[lo:hi] ** 2).sum(axis=1) + 100.0 * rank * 0.0)
        return gdist.all_gather_rows(local, N).numpy()

    sweeps = gdist.DrivenSweeps(sharded, dim=3, device="cpu")
    if rank != 0:
        q.put(("served", sweeps.serve()))
    else:
        rng = np.random.default_rng(os.getpid())  # only rank 0 draws: the ranks need not be in lock step
        ok = True
        try:
            with sweeps:  # releases the serving rank on exit, also when the driver raises
                for N in (7, 1, 12, 5):  # the particle count changes from sweep to sweep (rows outside the box are dropped)
                    rows = rng.standard_normal((N, 3))
                    ok = ok and np.allclose(sweeps(rows), (rows ** 2).sum(axis=1), rtol=0, atol=1e-15)
                raise RuntimeError("sampler failed")
        except RuntimeError:
            pass
        q.put(("driver", ok))
    dist.destroy_process_group()


def test_driven_sweeps_world2_gloo():
    """The multi-rank form of the SMC / MH particle sweep: rank 0 runs the sampler, rank 1 serves."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_driven_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["driver"] is True and res["served"] == 4


def test_bench_inputs_are_the_seeded_headline_case():
    sys.path.insert(0, ROOT)
    import bench
    from oracle import cases

    x, z, th = bench.headline_inputs(64, 8)
    xr, zr, thr = cases.headline(64, 8)
    assert np.array_equal(x, xr) and np.array_equal(z, zr) and np.array_equal(th, thr)
    assert bench.METRIC.startswith("REML logL+grad evals/s")


def test_bench_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-n", "1024"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["value"] > 0
    from oracle import vendor_ref

    assert line["cpu_baseline"]["kind"] == ("reference" if vendor_ref.available() else "port")
    assert line["cpu_baseline"]["cores"] == os.cpu_count()  # torchrun's OMP_NUM_THREADS=1 must not bite
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["metric"] == "REML logL+grad evals/s (n=8192,d=8,fp64)"


def test_build_entry_point():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    assert os.path.exists(os.path.join(ROOT, "gpmp_b200", "libgpmp_b200.so"))


# ---- drop-in binding (gpmp_b200/dropin.py) against the vendored reference, host logic only --------------------------
def test_dropin_install_patches_and_restores_the_reference():
    from oracle import vendor_ref

    if not vendor_ref.available() and not vendor_ref.vendor():
        pytest.skip("no reference tree here")
    gp = vendor_ref.import_reference("torch")
    import gpmp.core.model as model_mod
    import gpmp.kernel.matern as matern
    import gpmp.mcmc.param_posterior as post
    import gpmp.num as gnp
    from gpmp_b200 import dropin, kernel

    before = (model_mod.Model.predict, matern.maternp_covariance, gnp.scaled_distance, post.run_smc_sampling)
    dropin.install(gp)
    try:
        assert gp.kernel.maternp_covariance is kernel.maternp_covariance
        assert matern.maternp_covariance is kernel.maternp_covariance
        assert model_mod.Model.negative_log_restricted_likelihood.__name__ == "negative_log_restricted_likelihood"
        assert model_mod.Model.predict is not before[0] and post.run_smc_sampling is not before[3]
        dropin.install(gp)  # idempotent
    finally:
        dropin.uninstall()
    after = (model_mod.Model.predict, matern.maternp_covariance, gnp.scaled_distance, post.run_smc_sampling)
    assert after == before


def test_dropin_smc_uses_the_batched_sweep():
    """The reference's sample_from_selection_criterion_smc (param_posterior.py:658-775) with a BatchableCriterion:
    the wrapper around run_smc_sampling must find the criterion and the box inside the sampler's `logpdf_temp`
    closure and replace the per-particle loop; the target here is a closed-form quadratic (no device)."""
    from oracle import vendor_ref

    if not vendor_ref.available() and not vendor_ref.vendor():
        pytest.skip("no reference tree here")
    gp = vendor_ref.import_reference("torch")
    import gpmp.num as gnp
    from gpmp_b200 import dropin

    centre = np.array([0.5, -1.0])

    class Quad(dropin.BatchableCriterion):
        def __init__(self):
            self.sweeps, self.evaluations, self.scalar_calls, self.extra_term = 0, 0, 0, None

        def batched(self, thetas):
            th = np.asarray(thetas, dtype=np.float64).reshape(-1, 2)
            self.sweeps += 1
            self.evaluations += th.shape[0]
            return 0.5 * np.sum(((th - centre) / 0.3) ** 2, axis=1)

        def __call__(self, theta):
            self.scalar_calls += 1
            return float(self.batched(theta)[0])

    crit = Quad()
    box = [[-3.0, -4.0], [3.0, 3.0]]
    dropin.install(gp)
    try:
        particles, smc = gp.mcmc.sample_from_selection_criterion_smc(
            selection_criterion=crit, init_box=box, sampling_box=box, n_particles=300, mh_steps=5)
    finally:
        dropin.uninstall()
    P = gnp.to_np(particles)
    assert crit.scalar_calls == 0 and crit.sweeps > 5 and crit.evaluations > 300 * 5
    assert np.all(np.abs(P.mean(axis=0) - centre) <= 0.1) and np.all(np.abs(P.std(axis=0) - 0.3) <= 0.1)
    # the same call without the binding walks the particles one by one
    crit2 = Quad()
    gp.mcmc.sample_from_selection_criterion_smc(selection_criterion=crit2, init_box=box, sampling_box=box,
                                                n_particles=50, mh_steps=2)
    assert crit2.scalar_calls > 50


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
