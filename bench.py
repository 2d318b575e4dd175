#!/usr/bin/env python
"""Headline benchmark: REML logL + gradient evaluations per second at n=8192, d=8, fp64
(BASELINE.json configs[2]) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # CPU arm (oracle port of the reference)

One step = one evaluation of Model.negative_log_restricted_likelihood and its 9-component covparam
gradient (Matern p=2, constant mean) on x ~ U[0,1]^{8192x8}, z = sin(3 sum x) + 0.1 N(0,1), seed 1234.
A single evaluation does not shard ("replicas only"), so at N > 1 every rank evaluates its
own parameter vector (multi-start restarts): scaling is weak, value = N*K / max-over-ranks time.

value    : device-resident inputs (x, z already in HBM); theta (9 doubles) goes in by value and the
           scalar + gradient come back through pinned memory every step (the optimiser needs them).
e2e      : the same K steps through the public API with HOST numpy inputs: x, z and theta are copied
           host->device inside the timed region every step, value and gradient are read back.
roofline : FP64 tensor (DMMA) pipe.  achieved = n^3 flop per evaluation (potrf n^3/3 + trtri n^3/3 +
           lauum n^3/3) / step time -- a lower bound of the DMMA GEMM's rate, because its launches
           overlap on three streams; the per-launch CUDA-event sums (second pass over the same K steps, events
           on each launching stream) are reported per kernel class beside it.
"""
from __future__ import annotations

import argparse
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
FP64_NOMINAL_TFLOPS = 40.0  # HGX B200 datasheet; MEASURED_PEAKS.json carries no fp64 entry
# DRAM bytes of all DMMA launches (gemm_nt_kernel + trsm_tile_kernel) of one evaluation: dram__bytes_read.sum +
# dram__bytes_write.sum summed over the 201 launches, ncu capture committed as
# profiles/r01_gemm_dram_one_eval.csv (10.23 GB read + 1.29 GB written: 64x64 tiles re-read operands from L2/HBM
# more often than 128x128 tiles did -- 5.6 GB -- and still sit at ~0.5 TB/s, far below the HBM roofline)
GEMM_DRAM_BYTES_PER_STEP = 10.227e9 + 1.286e9


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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

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
def cpu_eval_seconds(n, d, threads_note=True):
    """One REML value + gradient on the host with the oracle port of the reference's torch backend
    (autograd through every op: gpmp/num/torch_backend.py:516-533 over core/likelihood.py:92-129)."""
    from oracle import gp_torch as ot

    x, z, th0 = headline_inputs(n, d)
    P = np.ones((n, 1))
    t0 = time.perf_counter()
    v, g = ot.reml_value_and_grad(x, z, P, P_MATERN, th0)
    return time.perf_counter() - t0, float(v)


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the reference is pure Python and
    cannot travel to the GPU box), all host threads, one bounded sample per step."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # full size when the run stays within a few minutes (one evaluation is ~7-25 s of host time),
    # otherwise n = 4096 scaled by n^3 (the O(n^2 d) terms make the scaled figure slightly pessimistic)
    n_s = args.ref_n if args.ref_n else (N_OBS if (args.steps + args.warmup) <= 8 else 4096)
    scale = (N_OBS / n_s) ** 3
    for _ in range(args.warmup):
        cpu_eval_seconds(n_s, DIM)
    ts = [cpu_eval_seconds(n_s, DIM)[0] for _ in range(args.steps)]
    total = sum(ts)
    value = args.steps / (total * scale)
    sample = (f"each step = 1 REML value+grad at n={n_s},d={DIM} (torch-CPU autograd, oracle port)"
              + ("" if n_s == N_OBS else f", time scaled by (8192/{n_s})^3={scale:.0f} to the n=8192 workload"))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total * scale / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "REML value+grad n=8192 d=8 Matern p=2 constant mean (configs[2])",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpus": os.cpu_count(),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
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

    # roofline pass: per-launch CUDA events on the launching stream over the same K steps
    _abi.prof_enable(True)
    torch.cuda.synchronize()
    for th in ths[args.warmup:]:
        step(th, xd, zd)
    torch.cuda.synchronize()
    prof = {}
    for cls, name in enumerate(["matern_cov", "dmma_gemm", "potf2", "dk_contract", "small", "batched"]):
        pms, cnt, work = _abi.prof_read(cls)
        prof[name] = {"ms_per_step": pms / args.steps, "launches_per_step": cnt / args.steps}
    _abi.prof_enable(False)
    flops = float(N_OBS) ** 3
    gemm_ms = prof["dmma_gemm"]["ms_per_step"]
    gemm_event_sum_tflops = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # The DMMA GEMM launches of one step overlap on three streams (bulk / chain / helper), so their event times
    # sum to MORE than the step; n^3 / step time is therefore a lower bound of the kernel's rate and is what is
    # reported as `achieved` (the event-sum figure is kept beside it).
    achieved = flops / (ms * 1e-3 / args.steps) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # cuBLAS dgemm on this GPU, for context next to the nominal peak
    A = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    for _ in range(2):
        A @ A
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        A @ A
    e1.record()
    torch.cuda.synchronize()
    cublas_tf = 5 * 2 * 4096**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del A

    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s",
        "frac": achieved / FP64_NOMINAL_TFLOPS, "traffic": GEMM_DRAM_BYTES_PER_STEP,
        "traffic_note": "bytes per step over all launches of the kernel (profiles/r01_gemm_dram_one_eval.csv)",
        "kernel": "gpmp::gemm_nt_kernel (FP64 DMMA.8x8x4)",
        "peak_source": "nominal HGX B200 FP64 (MEASURED_PEAKS.json has no fp64 entry); cuBLAS dgemm 4096^3 "
                       f"measured in this run: {cublas_tf:.1f} TFLOP/s",
        "algorithmic_flops_per_step": flops,
        "achieved_basis": "n^3 flop / step time: lower bound of the kernel's rate (its launches overlap on three "
                          "streams, so per-launch CUDA-event times sum to more than the step)",
        "gemm_event_sum_tflops": gemm_event_sum_tflops,
        "per_class": prof,
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        t_probe, _ = cpu_eval_seconds(2048, DIM)
        est = t_probe * 64
        n_s = N_OBS if est <= 45.0 else 4096
        t_s, _ = cpu_eval_seconds(n_s, DIM)
        scale = (N_OBS / n_s) ** 3
        cpu = {"value": 1.0 / (t_s * scale), "unit": "evals/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 REML value+grad at n={n_s},d={DIM} with the oracle port (torch-CPU autograd), "
                         + ("measured at full size" if n_s == N_OBS else f"scaled by (8192/{n_s})^3"),
               "host_cpus": os.cpu_count()}

    line = {
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "REML value+grad n=8192 d=8 Matern p=2 constant mean (BASELINE configs[2])",
                   "parallelism": f"replicas x{world} (one evaluation does not shard; independent restarts)",
                   "l2": "working set per step ~2.2 GB (L, T, T^T, K^-1) >> 126 MB L2; no flush needed",
                   "roofline_pass": "second pass over the same steps with per-launch CUDA events"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "last_value": last[0],
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpmp_b200", choices=["gpmp_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-n", type=int, default=0, help="reference arm: sample size override (tests)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
