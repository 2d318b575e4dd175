"""The CPU arm: one REML value + covparam gradient with the reference's own code.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/): used by bench.py (`--impl reference`, the
`cpu_baseline` leg and its same-theta parity check), tests/ and __graft_entry__.smoke().

`kind == "reference"`: the UNMODIFIED GPmp package vendored under oracle/_ref (oracle/vendor_ref.py), torch backend
on the CPU, exactly the call the reference's optimiser makes: gnp.value_and_grad (gpmp/num/torch_backend.py:516-533)
around Model.negative_log_restricted_likelihood (gpmp/core/model.py:405-427 -> core/likelihood.py:92-129) with
gp.kernel.maternp_covariance (gpmp/kernel/matern.py:124-141) and a constant mean.
`kind == "port"`: the restatement in oracle/gp_torch.py, only when oracle/_ref is absent.
"""
from __future__ import annotations

import os
import time

import numpy as np


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core."""
    import torch

    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


class RemlEvaluator:
    """Callable theta -> (value, grad[1+d], seconds) on fixed (x, z) with Matern p and a constant mean."""

    def __init__(self, x, z, p=2):
        from . import vendor_ref

        self.x, self.z, self.p = np.asarray(x, dtype=np.float64), np.asarray(z, dtype=np.float64), int(p)
        self.kind = "port"
        self._eval = None
        try:
            gp = vendor_ref.import_reference("torch")
        except ImportError:
            gp = None
        if gp is not None:
            import gpmp.num as gnp

            p_ = self.p
            model = gp.core.Model(lambda x_, param: gnp.ones((x_.shape[0], 1)),
                                  lambda a, b, cp, pairwise=False: gp.kernel.maternp_covariance(a, b, p_, cp, pairwise),
                                  None, None)
            xi, zi = gnp.asarray(self.x), gnp.asarray(self.z)

            def crit(theta):
                return model.negative_log_restricted_likelihood(theta, xi, zi)

            def ev(theta):
                v, g = gnp.value_and_grad(crit, gnp.asarray(np.asarray(theta, dtype=np.float64)))
                return float(v), np.asarray(g.detach().cpu().numpy(), dtype=np.float64)

            self._eval, self.kind = ev, "reference"
        else:
            from . import gp_torch as ot

            P = np.ones((self.x.shape[0], 1))

            def ev(theta):
                v, g = ot.reml_value_and_grad(self.x, self.z, P, self.p, np.asarray(theta, dtype=np.float64))
                return float(v), np.asarray(g, dtype=np.float64)

            self._eval = ev

    def __call__(self, theta):
        t0 = time.perf_counter()
        v, g = self._eval(theta)
        return v, g, time.perf_counter() - t0

    def describe(self):
        if self.kind == "reference":
            return ("GPmp 0.9.37 itself (oracle/_ref, torch backend on the CPU): gnp.value_and_grad of "
                    "Model.negative_log_restricted_likelihood")
        return "oracle port of the reference's torch backend (oracle/gp_torch.py; oracle/_ref not present)"
