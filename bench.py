#!/usr/bin/env python
"""Headline benchmark: REML logL + gradient evaluations per second at n=8192, d=8, fp64
(BASELINE.json configs[2]) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm: GPmp itself (oracle/_ref) on the host

One step = one evaluation of Model.negative_log_restricted_likelihood and its 9-component covparam
gradient (Matern p=2, constant mean) on x ~ U[0,1]^{8192x8}, z = sin(3 sum x) + 0.1 N(0,1), seed 1234.
A single evaluation does not shard ("replicas only"), so at N > 1 every rank evaluates its
own parameter vector (multi-start restarts): scaling is weak, value = N*K / max-over-ranks time.

value     : device-resident inputs (x, z already in HBM); theta (9 doubles) goes in by value and the
            scalar + gradient come back through pinned memory every step (the optimiser needs them).
e2e       : the same K steps through the public API with HOST numpy inputs: x, z and theta are copied
            host->device inside the timed region every step, value and gradient are read back.
roofline  : FP64 tensor (DMMA) pipe.  achieved = n^3 flop per evaluation (potrf n^3/3 + trtri n^3/3 +
            lauum n^3/3) / step time.  peak = the DMMA issue rate MEASURED in this run
            (gpmp_measure_dmma_peak: every SM issuing DMMA.8x8x4 from registers), with cuBLAS dgemm 8192^3
            (best of 10) and the nominal 40 TFLOP/s beside it.  per_class: CUDA-event times per kernel class
            from a serialised pass (all look-ahead streams folded onto one stream: exclusive times) and from
            the overlapped production schedule, with the algorithmic work of each class.
parity    : the CPU arm evaluates the SAME theta as the last timed GPU step; value / gradient relative
            errors are printed and the bench fails above 1e-8 (BASELINE.json tolerance).
secondary : the paths north_star wants to SCALE, timed across the same N ranks: the 8192-particle x n=512
            batched REML sweep (rows sharded, one all-gather) and the panel-partitioned n=32768 value + gradient
            (--no-secondary / --no-secondary-large skip them).
"""
from __future__ import annotations

import argparse
import csv
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM, P_MATERN, SEED = 8192, 8, 2, 1234
METRIC = "REML logL+grad evals/s (n=8192,d=8,fp64)"
WORKLOAD = "REML value+grad n=8192 d=8 Matern p=2 constant mean (BASELINE configs[2])"
FP64_NOMINAL_TFLOPS = 40.0  # HGX B200 datasheet; MEASURED_PEAKS.json carries no fp64 entry
TOL_PARITY = 1e-8
# ncu capture of the DRAM traffic of every DMMA launch of one evaluation (dram__bytes_read/write.sum per launch);
# summed at run time so the line always quotes the committed capture it names
TRAFFIC_CSVS = ["profiles/r02_gemm_dram_one_eval.csv", "profiles/r01_gemm_dram_one_eval.csv"]


def headline_inputs(n=N_OBS, d=DIM, seed=SEED):
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, d))
    z = np.sin(3.0 * x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    th0 = np.concatenate(([0.0], np.full(d, -np.log(0.7))))
    return x, z, th0


def thetas_for(th0, count, rank):
    rng = np.random.default_rng(1000 + rank)
    return [th0 + rng.uniform(-0.25, 0.25, size=th0.shape) for _ in range(count)]


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "samples": len(sm), "reasons": sorted(reasons)}
        if not sm and self.lines:
            out["raw"] = self.lines[0][:200]
        return out


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: GPmp's own CPU implementation of the step (oracle/_ref: the unmodified package, torch
    backend, gnp.value_and_grad of Model.negative_log_restricted_likelihood) on all host cores, at the FULL
    n=8192 workload.  One evaluation is ~10 s of host time, so at most 3 evaluations are timed (and at most one
    warmed up) whatever --steps / --warmup ask for; the line says so.  Under torchrun rank 0 alone runs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_eval

    cores = ref_eval.use_all_host_threads()
    n_s = args.ref_n if args.ref_n else N_OBS
    x, z, th0 = headline_inputs(n_s, DIM)
    ev = ref_eval.RemlEvaluator(x, z, P_MATERN)
    timed = max(1, min(args.steps, 3))
    warm = min(args.warmup, 1)
    ths = thetas_for(th0, warm + timed, 0)
    for th in ths[:warm]:
        ev(th)
    ts = [ev(th)[2] for th in ths[warm:]]
    total = sum(ts)
    value = timed / total
    sample = (f"{timed} full-size evaluations timed (of --steps {args.steps}; {warm} warm-up): each = 1 REML value + "
              f"gradient at n={n_s}, d={DIM}, by {ev.describe()}, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": timed, "warmup": warm, "steps_requested": args.steps, "ms_per_step": 1e3 * total / timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD if n_s == N_OBS else f"REML value+grad n={n_s} (test override)",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": ev.kind, "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def measure_fp64_peaks(torch, _abi):
    """(DMMA issue-rate TFLOP/s, cuBLAS dgemm 8192^3 TFLOP/s), both best of 10 with CUDA events."""
    import ctypes as C

    lib = _abi.lib()
    sink = torch.zeros(148 * 8 * 512, dtype=torch.float64, device="cuda")
    flops = C.c_double()
    best = 0.0
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.check(lib.gpmp_measure_dmma_peak(148 * 8, 20000, _abi.ptr(sink), C.byref(flops), _abi.stream_ptr()),
                   "gpmp_measure_dmma_peak")
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    A = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    B = torch.empty_like(A)
    cub = 0.0
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(A, A, out=B)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            cub = max(cub, 2 * 8192**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del A, B
    return best, cub


def committed_traffic():
    """DRAM bytes per evaluation over all DMMA launches, summed from the committed ncu capture (or None)."""
    for rel in TRAFFIC_CSVS:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        total, hit = 0.0, False
        with open(path, newline="") as f:
            rows = [r for r in csv.reader(f) if r]
        header = next((r for r in rows if "Metric Name" in r and "Metric Value" in r), None)
        if header is None:
            continue
        iname, ival, iunit = header.index("Metric Name"), header.index("Metric Value"), header.index("Metric Unit")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows:
            if len(r) <= max(iname, ival, iunit) or not r[iname].startswith("dram__bytes_"):
                continue
            try:
                total += float(r[ival].replace(",", "")) * scale.get(r[iunit], 1.0)
                hit = True
            except ValueError:
                pass
        if hit:
            return total, rel
    return None, None


def secondary_paths(args, torch, dist, gp, world, rank, peak_tf):
    """The sharded paths of SURVEY.md 8(e), timed on the same ranks with the same barrier / max-over-ranks rule."""
    out = {}
    group = dist.group.WORLD if world > 1 else None

    def timed(fn, reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / reps

    # --- BASELINE configs[3]: 8192 particles x (n=512, d=4), REML values, rows sharded over the ranks
    n, d, N = 512, 4, 8192
    rng = np.random.default_rng(4321)
    x = rng.uniform(size=(n, d))
    z = np.sin(3.0 * x.sum(axis=1)) + 0.1 * rng.standard_normal(n)
    th_hat = np.concatenate(([0.0], np.full(d, -np.log(0.5))))
    TH = th_hat + rng.uniform(-2.0, 2.0, size=(N, d + 1))
    model = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                          lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, P_MATERN, cp, pairwise))
    crit = gp.batched.BatchedCriterion(model, x, z, P_MATERN, kind="reml", group=group)
    vals = crit(TH)
    crit(TH)
    ms = timed(lambda: crit(TH), 5)
    flops = N * float(n) ** 3 / 3.0
    tf = flops / (ms * 1e-3) / 1e12
    out["smc_sweep"] = {
        "workload": "8192 particles x REML value at n=512, d=4 (BASELINE configs[3]); host theta in, host values out",
        "ms_per_sweep": ms, "sweeps_per_s": 1e3 / ms, "particle_evals_per_s": N * 1e3 / ms,
        "tflops_aggregate": tf, "frac_of_peak_per_gpu": tf / world / peak_tf,
        "algorithmic_flops_per_sweep": flops, "finite_values": int(np.isfinite(vals).sum()),
        "sharding": f"theta rows block-partitioned over {world} rank(s); one all-gather of N values per sweep",
    }
    if not args.no_secondary_large:
        # --- BASELINE configs[4]: one n=32768, d=10 REML value + gradient, factorisation / inverse partitioned
        n2, d2 = 32768, 10
        x2, z2, _ = headline_inputs(n2, d2, seed=99)
        th2 = np.concatenate(([0.0], np.full(d2, -np.log(0.7))))
        m2 = gp.core.Model(lambda x_, mp: gp.num.ones((x_.shape[0], 1)),
                           lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, P_MATERN, cp, pairwise))
        xd, zd = gp.num.asarray(x2), gp.num.asarray(z2)

        def partitioned():
            gp.dist.reml_value_and_grad_distributed(m2, th2, xd, zd, group)

        def single():  # the plain one-GPU path (look-ahead Cholesky, block-doubling inverse): the honest N = 1 point
            tp = torch.tensor(th2, requires_grad=True)
            v = m2.negative_log_restricted_likelihood(tp, xd, zd)
            torch.autograd.grad(v, tp)

        fn = single if world == 1 else partitioned
        fn()
        ms2 = timed(fn, 2)
        tf2 = float(n2) ** 3 / (ms2 * 1e-3) / 1e12
        out["partitioned_reml_n32768"] = {
            "workload": "one REML value+grad at n=32768, d=10 (BASELINE configs[4]); panel-partitioned over the ranks"
                        + (" (one rank: the single-GPU path, which is faster than the partitioned code on one rank)"
                           if world == 1 else ""),
            "ms_per_eval": ms2, "tflops_aggregate": tf2, "frac_of_peak_per_gpu": tf2 / world / peak_tf,
            "algorithmic_flops": float(n2) ** 3,
        }
        if world == 1:
            partitioned()
            out["partitioned_reml_n32768"]["ms_per_eval_partitioned_code_one_rank"] = timed(partitioned, 2)
        del xd, zd
        torch.cuda.empty_cache()
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import gpmp_b200 as gp
    from gpmp_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _abi.lib()

    x, z, th0 = headline_inputs()
    gnp = gp.num
    xd, zd = gnp.asarray(x), gnp.asarray(z)
    model = gp.core.Model(lambda x_, mp: gnp.ones((x_.shape[0], 1)),
                          lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, P_MATERN, cp, pairwise))

    def step(theta, xin, zin):
        tp = torch.tensor(theta, requires_grad=True)
        v = model.negative_log_restricted_likelihood(tp, xin, zin)
        (g,) = torch.autograd.grad(v, tp)
        return v.item(), g.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(thetas, xin, zin):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for th in thetas:
            last = step(th, xin, zin)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item()), last

    ths = thetas_for(th0, args.steps + args.warmup, rank)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # sampled under load: warm-up, the timed region and the e2e pass
    for th in ths[: args.warmup]:
        step(th, xd, zd)
    l0 = _abi.launch_count()
    ms, last = timed(ths[args.warmup:], xd, zd)
    launches = _abi.launch_count() - l0
    value = world * args.steps / (ms * 1e-3)

    # e2e: HOST inputs every step (x, z in pinned host memory, theta a host vector; all three are copied
    # host->device inside the timed region, value and gradient are read back)
    xh, zh = torch.from_numpy(x).pin_memory(), torch.from_numpy(z).pin_memory()
    for th in ths[:2]:
        step(th, xh, zh)
    ms_e2e, _ = timed(ths[args.warmup:], xh, zh)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    clocks = sampler.stop() if rank == 0 else None
    h2d = 8 * (x.size + z.size + th0.size)
    d2h = 8 * 8 + 8 * th0.size  # 64-byte value/info record + gradient

    # roofline passes over (up to 5 of) the same steps with per-launch CUDA events: overlapped production
    # schedule (times overlap across the look-ahead streams) and serialised (exclusive times)
    names = ["matern_cov", "dmma_gemm", "potf2", "dk_contract", "small", "batched"]
    prof_steps = ths[args.warmup:][: min(5, args.steps)]
    prof = {}
    for mode, key in ((2, "serialised"), (1, "overlapped")):
        _abi.prof_enable(mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for th in prof_steps:
            step(th, xd, zd)
        e1.record()
        torch.cuda.synchronize()
        per = {}
        for cls, name in enumerate(names):
            pms, cnt, work = _abi.prof_read(cls)
            k = len(prof_steps)
            per[name] = {"ms_per_step": pms / k, "launches_per_step": cnt / k, "work_per_step": work / k}
            if name in ("dmma_gemm", "potf2") and pms > 0:
                per[name]["tflops"] = work / (pms * 1e-3) / 1e12
            if name in ("matern_cov", "dk_contract") and pms > 0:
                per[name]["gbytes_per_s"] = work / (pms * 1e-3) / 1e9
        _abi.prof_enable(0)
        prof[key] = {"step_ms_with_events": e0.elapsed_time(e1) / len(prof_steps), "per_class": per}
    flops = float(N_OBS) ** 3
    achieved = flops / (ms * 1e-3 / args.steps) / 1e12

    dmma_peak, cublas_tf = measure_fp64_peaks(torch, _abi)
    secondary = None
    if not args.no_secondary:
        secondary = secondary_paths(args, torch, dist, gp, world, rank, dmma_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic, traffic_src = committed_traffic()
    ser = prof["serialised"]["per_class"]
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
        "frac": achieved / dmma_peak, "traffic": traffic,
        "traffic_note": (f"dram__bytes_read+write summed over all DMMA launches of one evaluation ({traffic_src})"
                         if traffic_src else "no committed ncu capture found"),
        "kernel": "gpmp::gemm_nt_tma_kernel + gpmp::gemm_nt_kernel (FP64 DMMA.8x8x4; the TMA / mbarrier 128x128 kernel for the long-k products, the cp.async 64x64 kernel for the K = 512 updates)",
        "peak_source": "FP64 tensor pipe issue rate measured in this run (gpmp_measure_dmma_peak: 148x8 CTAs x 16 warps "
                       "x 8 independent DMMA.8x8x4 chains from registers, best of 10); MEASURED_PEAKS.json has no fp64 "
                       "entry",
        "peak_cublas_dgemm_8192": cublas_tf, "peak_nominal": FP64_NOMINAL_TFLOPS,
        "frac_of_cublas_dgemm": achieved / cublas_tf, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
        "algorithmic_flops_per_step": flops,
        "achieved_basis": "n^3 flop / step time (whole step: K build, factorisation, inverse, contraction, readbacks)",
        "kernel_rate_exclusive_tflops": ser["dmma_gemm"].get("tflops"),
        "kernel_share_of_serialised_step": (ser["dmma_gemm"]["ms_per_step"]
                                            / max(1e-9, sum(c["ms_per_step"] for c in ser.values()))),
        "passes": prof,
    }

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import ref_eval

        cores = ref_eval.use_all_host_threads()
        ev = ref_eval.RemlEvaluator(x, z, P_MATERN)
        th_last = ths[-1]
        if args.cpu_warm:
            ev(th_last)
        v_ref, g_ref, t_s = ev(th_last)
        cpu = {"value": 1.0 / t_s, "unit": "evals/s", "cores": cores, "kind": ev.kind,
               "sample": f"1 full-size evaluation (n={N_OBS}, d={DIM}) at the theta of the last timed GPU step, by "
                         f"{ev.describe()}, {cores} threads, {t_s:.1f} s",
               "host_cpus": os.cpu_count()}
        v_gpu, g_gpu = last
        parity = {"theta": "last timed step", "value_gpu": v_gpu, "value_cpu": v_ref,
                  "value_rel": abs(v_gpu - v_ref) / abs(v_ref),
                  "grad_rel": float(np.max(np.abs(g_gpu - g_ref)) / np.max(np.abs(g_ref))),
                  "tolerance": TOL_PARITY}
        parity["ok"] = bool(parity["value_rel"] <= TOL_PARITY and parity["grad_rel"] <= TOL_PARITY)

    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "parallelism": f"replicas x{world} (one evaluation does not shard; independent restarts)",
                   "l2": "working set per step ~2.2 GB (L, T, T^T, K^-1) >> 126 MB L2; no flush needed",
                   "roofline_pass": "two extra passes over the same steps with per-launch CUDA events"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "secondary": secondary,
        "last_value": last[0],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print(f"bench: parity check failed: {parity}", file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpmp_b200", choices=["gpmp_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-warm", action="store_true", help="cpu_baseline: one untimed evaluation first")
    ap.add_argument("--no-secondary", action="store_true", help="skip the sharded config-4 sweep")
    ap.add_argument("--no-secondary-large", action="store_true",
                    help="skip the panel-partitioned n=32768 value+grad of the secondary block")
    ap.add_argument("--ref-n", type=int, default=0, help="reference arm: sample size override (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
