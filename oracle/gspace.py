"""Oracle restatement (TEST INFRASTRUCTURE) of the BASQ kernel handed to ``recombination`` by ``BASQ.quadrature``
(SOBER/BASQ/_basq.py:55-67): ``ScaleMmltGP.gspace_kernel`` (SOBER/BASQ/_scale_mmlt.py:256-275) and the predictions it
is built from (:208-220).  SURVEY.md 8(f) row 4.

    mu_g(x)    = exp(mu_h(x) + var_h(x) / 2) - 1                     (mu_h, var_h) = predict(x, model)
    k_g(x, y)  = mu_g(x) mu_g(y) (exp(cov_h(x, y)) - 1)              cov_h = predictive_covariance(x, y, model)
    + jitter on the entries [i, i], i < min(len(x), len(y))          (jitter = 0 in the reference's constructor, :70)

``predict`` / ``predictive_covariance`` are the gpytorch-free restatements of oracle/gp.py and oracle/kernels.py; the
class below is pinned against the reference's own methods by tests/golden/make_golden_gspace.py (which loads the
unmodified _scale_mmlt.py and binds those two functions into its ``_gp`` import).
"""
import torch

from . import gp as ogp
from . import kernels as ok


class ScaleMmltGP:
    """The part of SOBER/BASQ/_scale_mmlt.py a kernel callable needs: ``model`` (a fitted GP in h space) and ``jitter``."""

    def __init__(self, model, jitter=0.0):
        self.model = model
        self.jitter = torch.as_tensor(jitter, dtype=torch.float64)

    def gspace_mean_predict(self, x):
        mu_h, var_h = ogp.predict(x, self.model)
        return (mu_h + 0.5 * var_h).exp() - 1

    def hspace_kernel(self, x, y):
        return ok.predictive_covariance(x, y, self.model)

    def gspace_kernel(self, x, y):
        mu_g_x = self.gspace_mean_predict(x)
        mu_g_y = self.gspace_mean_predict(y)
        cov_h_xy = self.hspace_kernel(x, y)
        if len(cov_h_xy.shape) == 2:
            out = mu_g_x.unsqueeze(1) * mu_g_y.unsqueeze(0) * (cov_h_xy.exp() - 1)
        elif len(cov_h_xy.shape) == 3:
            out = mu_g_x.unsqueeze(1).unsqueeze(0) * mu_g_y.unsqueeze(1) * (cov_h_xy.exp() - 1)
        d = min(len(x), len(y))
        out[range(d), range(d)] = out[range(d), range(d)] + self.jitter
        return out
