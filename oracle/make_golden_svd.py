"""Golden vectors for sample_paths(method="svd") (gpmp/core/sample_paths.py:50-58: the symmetric square root
U sqrt(s) U^T of K(xt, xt), meant for covariance matrices Cholesky cannot factor), produced by running the
reference's own lines on the UNMODIFIED GPmp 0.9.37 (/root/reference, numpy backend) in this container and committed
as tests/golden/reference_svd.npz.  TEST INFRASTRUCTURE ONLY.

    GPMP_BACKEND=numpy python oracle/make_golden_svd.py

The reference draws its normals inside sample_paths (unseedable from outside), so -- as for the Cholesky route -- the
fixture holds the deterministic map: C = U sqrt(s) V^T from gnp.svd(K, hermitian=True) and C @ normals for fixed
normals.  Cases: duplicated points (K exactly singular: the Cholesky route fails), a dense 1-d design
(cond ~ 1e17), and a well-conditioned matrix (both routes work).
"""
import os
import sys

import numpy as np

os.environ.setdefault("GPMP_BACKEND", "numpy")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

CASES = {  # name: (n distinct points, duplicated points appended, d, p, seed)
    "svd_dup_n200_d2_p2": (150, 50, 2, 2, 11),
    "svd_dense_n300_d1_p2": (300, 0, 1, 2, 12),
    "svd_pd_n120_d3_p1": (120, 0, 3, 1, 13),
}


def inputs(name):
    n, dup, d, p, seed = CASES[name]
    rng = np.random.default_rng(seed)
    x = rng.uniform(size=(n, d))
    if dup:
        x = np.vstack([x, x[:dup]])
    th = np.concatenate(([0.3], np.full(d, 0.5)))
    normals = rng.standard_normal((x.shape[0], 7))
    return x, th, p, normals


def main():
    from oracle import vendor_ref

    vendor_ref.ensure_plot_stub()
    import gpmp as gp
    import gpmp.num as gnp

    out = {}
    for name in CASES:
        x, th, p, normals = inputs(name)
        model = gp.core.Model(lambda x_, mp: gnp.ones((x_.shape[0], 1)),
                              lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p, cp, pairwise),
                              None, gnp.asarray(th))
        xt_ = gnp.asarray(x)
        K = model.covariance(xt_, xt_, model.covparam)
        # gpmp/core/sample_paths.py:50-55
        U, s, Vt = gnp.svd(K, full_matrices=True, hermitian=True)
        C = gnp.matmul(U * gnp.sqrt(s), Vt)
        zsim = gnp.matmul(C, gnp.asarray(normals))
        out[name] = dict(x=x, theta=th, p=np.array(p), normals=normals,
                         C=np.asarray(gnp.to_np(C)), zsim=np.asarray(gnp.to_np(zsim)), s=np.asarray(gnp.to_np(s)))
        print(name, "cond", float(s[0] / max(s[-1], 1e-300)))
    flat = {f"{c}/{k}": np.asarray(v) for c, rec in out.items() for k, v in rec.items()}
    path = os.path.join(ROOT, "tests", "golden", "reference_svd.npz")
    np.savez_compressed(path, **flat)
    print(f"wrote {path}: {len(flat)} arrays, {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
