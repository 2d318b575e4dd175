"""gpmp_b200: the exact-GP inner loop of GPmp (gpmp.kernel / gpmp.core / the hot slice of gpmp.num) on B200.

    import gpmp_b200 as gp
    gnp = gp.num
    model = gp.core.Model(mean, covariance, meanparam=None, covparam=theta)
    nlrl = model.negative_log_restricted_likelihood(theta, xi, zi)     # 0-d tensor with grad_fn
    zpm, zpv = model.predict(xi, zi, xt)

All arithmetic runs in libgpmp_b200.so (hand-written sm_100a kernels behind the C-ABI of
include/gpmp_b200.h).  There is no CPU fallback: importing works anywhere, calling needs a CUDA device
and the built library.
"""
from . import _abi, batched, core, dist, fisher, kernel, num, ops, selection  # noqa: F401
from .core import Model  # noqa: F401

__version__ = "0.1.0"
