"""Development: latency of the chain's small launches in isolation: 50 dependent launches captured in a CUDA
graph (no host overhead), time per launch."""
import sys
import torch
sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import ops

torch.manual_seed(0)
dev = ops.device()
def graph_time(fn, reps=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (M, N, K, lower) in [(128, 128, 128, True), (128, 128, 256, True), (128, 128, 512, True), (2048, 128, 128, False),
                         (4096, 128, 128, False), (8192, 128, 128, False), (64, 64, 16, False), (64, 64, 128, False),
                         (64, 64, 1024, False)]:
    A = ops.padded(torch.randn(M, K, dtype=torch.float64, device=dev))
    B = ops.padded(torch.randn(N, K, dtype=torch.float64, device=dev))
    Cm = ops.padded(torch.zeros(M, N, dtype=torch.float64, device=dev))
    us = graph_time(lambda: ops.gemm_nt(A, B, C_out=Cm, alpha=-1.0, beta=1.0, lower=lower))
    print(f"gemm_nt M={M} N={N} K={K} lower={lower}: {us:.1f} us per launch")
x = torch.zeros(1024, device=dev)
print("tiny torch kernel:", round(graph_time(lambda: x.add_(1.0)), 2), "us")
