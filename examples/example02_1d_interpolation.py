"""BASELINE config 1 with gpmp_b200: 1-D interpolation, Matern p=3, constant mean, REML selection, prediction
at 200 points (the scenario of GPmp's examples/gpmp_example02_1d_interpolation.py, without the plots)."""
import numpy as np

import gpmp_b200 as gp

gnp = gp.num


def twobumps(x):
    # gpmp/misc/testfunctions.py: two Gaussian bumps on [-1, 1]
    x = np.asarray(x).reshape(-1)
    return 0.7 * np.exp(-((x + 0.4) ** 2) / 0.02) + np.exp(-((x - 0.3) ** 2) / 0.05)


def main():
    rng = np.random.default_rng(1234)
    xi = np.sort(rng.uniform(-1.0, 1.0, size=(6, 1)), axis=0)
    zi = twobumps(xi)
    xt = np.linspace(-1.0, 1.0, 200).reshape(-1, 1)
    p = 3
    model = gp.core.Model(lambda x, meanparam: gnp.ones((x.shape[0], 1)),
                          lambda x, y, covparam, pairwise=False: gp.kernel.maternp_covariance(x, y, p, covparam, pairwise))
    model, info = gp.kernel.select_parameters_with_reml(model, xi, zi, info=True)
    zpm, zpv = model.predict(xi, zi, xt)
    print("covparam0 :", info["covparam0"])
    print("covparam  :", np.asarray(model.covparam), " criterion:", float(info.fun), " iterations:", info.nit)
    print("max |error| on the grid:", float(np.max(np.abs(zpm - twobumps(xt)))),
          " max predictive sd:", float(np.sqrt(zpv.max())))
    zloo, s2loo, eloo = model.loo(xi, zi, convert_out=True)
    print("LOO errors:", eloo)


if __name__ == "__main__":
    main()
