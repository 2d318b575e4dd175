"""Golden vectors for the Fisher information (tests/golden/reference_fisher.npz), produced by the UNMODIFIED
reference's `gpmp.core.fisher` (NumPy backend) in this container:

    GPMP_BACKEND=numpy python oracle/make_golden_fisher.py

Test infrastructure only; never run on the GPU box (the reference does not travel).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

# name, n, d, p, mean kind, noisy kernel, seed
FISHER_CASES = [
    ("fisher_n120_d3_p2_const", 120, 3, 2, "const", False, 41),
    ("fisher_n90_d2_p1_zero", 90, 2, 1, "zero", False, 42),
    ("fisher_n100_d2_p2_linear_noise", 100, 2, 2, "linear", True, 43),
]


def main():
    os.environ.setdefault("GPMP_BACKEND", "numpy")
    from oracle import cases
    import gpmp as gp
    import gpmp.num as gnp
    from gpmp.core import fisher

    def cov_fn(p, noise):
        if not noise:
            return lambda x, y, covparam, pairwise=False: gp.kernel.maternp_covariance(x, y, p, covparam, pairwise)

        def k(x, y, param, pairwise=False):  # examples/gpmp_example07_nd_regression.py:95-130
            sigma2, loginvrho = gnp.exp(param[0]), param[2:]
            if y is x or y is None:
                if pairwise:
                    return sigma2 * gnp.ones((x.shape[0],))
                K = gnp.scaled_distance(loginvrho, x, x)
                return sigma2 * gp.kernel.maternp_kernel(p, K) + gnp.exp(param[1]) * gnp.eye(K.shape[0])
            K = gnp.scaled_distance_elementwise(loginvrho, x, y) if pairwise else gnp.scaled_distance(loginvrho, x, y)
            return sigma2 * gp.kernel.maternp_kernel(p, K)
        return k

    out = {}
    for name, n, d, p, kind, noise, seed in FISHER_CASES:
        x, _, _ = cases.data(n, d, seed)
        th = cases.theta(d, seed, noise=noise)
        model = gp.core.Model(cases.mean_fn(kind, gnp), cov_fn(p, noise), meanparam=None, covparam=gnp.asarray(th),
                              meantype=cases.meantype_of(kind))
        xg = gnp.asarray(x)
        out[name + "/x"], out[name + "/theta"] = x, th
        out[name + "/spd"] = np.asarray(fisher.fisher_information(model, xg, gnp.asarray(th)))
        out[name + "/cpd"] = np.asarray(fisher.fisher_information_cpd(model, xg, gnp.asarray(th)))
    path = os.path.join(ROOT, "tests", "golden", "reference_fisher.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("pd")})


if __name__ == "__main__":
    main()
