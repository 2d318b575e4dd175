"""BASELINE config 1: GPmp's examples/gpmp_example02_1d_interpolation.py scenario (without the plots) with the
B200 hot path bound into GPmp itself: the model, the REML / REMAP selection drivers and the priors are GPmp's own
code; `gpmp_b200.dropin.install()` routes the kernel / Model seams to the CUDA kernels.

    GPMP_BACKEND=torch python examples/example02_1d_interpolation.py      # needs GPmp importable as `gpmp`
"""
import os

os.environ.setdefault("GPMP_BACKEND", "torch")
import numpy as np

try:
    import gpmp as gp
except ImportError:  # in this repository GPmp lives under oracle/_ref (test infrastructure, not shipped)
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import vendor_ref

    gp = vendor_ref.import_reference("torch")
import gpmp.num as gnp

import gpmp_b200.dropin as b200


def main():
    b200.install(gp)
    rng = np.random.default_rng(1234)
    xi = np.sort(rng.uniform(-1.0, 1.0, size=(6, 1)), axis=0)
    zi = gp.misc.testfunctions.twobumps(xi)
    xt = np.linspace(-1.0, 1.0, 200).reshape(-1, 1)
    p = 3

    def model():
        return gp.core.Model(lambda x, meanparam: gnp.ones((x.shape[0], 1)),
                             lambda x, y, covparam, pairwise=False: gp.kernel.maternp_covariance(x, y, p, covparam, pairwise))

    for name, select in (("REML", gp.kernel.select_parameters_with_reml), ("REMAP", gp.kernel.select_parameters_with_remap)):
        m, info = select(model(), xi, zi, info=True)
        zpm, zpv = m.predict(xi, zi, xt)
        print(f"{name}: covparam {gnp.to_np(m.covparam)}  criterion {float(info.fun):.6f}  iterations {info.nit}")
        print(f"      max |error| on the grid {float(np.max(np.abs(zpm - gp.misc.testfunctions.twobumps(xt)))):.4f}"
              f"  max predictive sd {float(np.sqrt(zpv.max())):.4f}")
    zloo, s2loo, eloo = m.loo(xi, zi, convert_out=True)
    print("LOO errors:", eloo)


if __name__ == "__main__":
    main()
