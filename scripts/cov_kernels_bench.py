"""Development: K1 (covariance build, lower-triangle mode) and K4 (dK contraction) in isolation, GB/s of
algorithmic bytes (8 n (n+1) / 2) for several input dimensions, plus the full-matrix public op."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import gpmp_b200 as gp
from gpmp_b200 import _abi, ops

lib = _abi.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
out = {}


def ev(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


rng = np.random.default_rng(0)
ld = ops._round_ld(n)
K = torch.empty((n, ld), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for d in (1, 2, 4, 8, 16):
    for p in (2,) if d != 8 else (0, 2, 4):
        x = gp.num.asarray(rng.uniform(size=(n, d)))
        spec = _abi.make_spec(p, d, 0.1, [-np.log(0.7)] * d)
        f1 = lambda: _abi.check(lib.gpmp_matern_cov(C.byref(spec), _abi.ptr(x), n, None, 0, _abi.ptr(K), ld,
                                                    _abi.COV_LOWER, _abi.stream_ptr()), "cov")
        t = ev(f1)
        byt = 8.0 * n * (n + 1) / 2
        rec = {"k1_lower_ms": t, "k1_lower_gbs": byt / t / 1e6}
        f2 = lambda: _abi.check(lib.gpmp_matern_cov(C.byref(spec), _abi.ptr(x), n, None, 0, _abi.ptr(K), ld,
                                                    _abi.COV_FULL, _abi.stream_ptr()), "cov")
        t = ev(f2)
        rec.update({"k1_full_ms": t, "k1_full_gbs": 8.0 * n * n / t / 1e6})
        # K4 through the public backward of the covariance op: G = K (any matrix will do), same-set rectangular form
        ws = torch.empty(lib.gpmp_contract_workspace_bytes(n, n, d), dtype=torch.uint8, device="cuda")
        g = torch.empty(1 + d, dtype=torch.float64, device="cuda")
        f3 = lambda: _abi.check(lib.gpmp_matern_cov_backward(C.byref(spec), _abi.ptr(x), n, None, 0, _abi.ptr(K), ld,
                                                             _abi.ptr(g), _abi.ptr(ws), ws.numel(),
                                                             _abi.stream_ptr()), "bwd")
        t = ev(f3)
        rec.update({"k4_rect_ms": t, "k4_rect_gbs": 8.0 * n * n / t / 1e6})
        out[f"d{d}_p{p}"] = rec
print(json.dumps(out, indent=1))
