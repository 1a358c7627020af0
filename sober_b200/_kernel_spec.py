"""Introspection of the reference's kernel objects into a plain description the CUDA path can fuse.

The reference hands ``recombination`` an opaque callable (``SOBER/_rchq.py:9``).  In every example it is a
``SOBER._kernel.Kernel(model, mode)`` (``SOBER/_kernel.py:4-30``) around a gpytorch model whose
``covar_module`` is ``ScaleKernel(RBFKernel | MaternKernel)`` or ``ScaleKernel(TanimotoKernel)``
(``SOBER/_drug_modelling.py:86-107``).  Those objects stay untouched; this module only *reads* them
(duck-typed on class names and attributes, so real gpytorch modules and the test stand-ins both work).

Anything that cannot be described returns ``None`` and the caller takes the generic path, where the callable
itself is evaluated tile by tile on the device.
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


@dataclass
class KernelSpec:
    family: int                      # _lib.RBF ... _lib.TANIMOTO
    d: Optional[int]                 # input dimension if the lengthscale pins it (ARD), else None
    inv_ls: Optional[torch.Tensor]   # (d,) or (1,) reciprocal lengthscales; None for Tanimoto
    outputscale: float
    mode: str                        # "kernel" | "predictive_covariance"
    x_obs: Optional[torch.Tensor] = None     # (n_obs, d) training inputs      (predictive covariance)
    woodbury: Optional[torch.Tensor] = None  # (n_obs, n_obs)  S S^T           (SOBER/_gp.py:255-278)
    alpha: Optional[torch.Tensor] = None     # (n_obs,) mean cache             (weighted mode: m(x) = c + k(x, X) alpha)
    mean_const: float = 0.0
    noise: float = 0.0                       # gspace mode: likelihood noise (enters the predictive variance)
    jitter: float = 0.0                      # gspace mode: ScaleMmltGP.jitter

    @property
    def gspace(self):
        return self.mode == "gspace"

    @property
    def weighted(self):
        return self.mode == "weighted_predictive_covariance"

    @property
    def posterior(self):
        """the covariance part is the GP posterior predictive covariance (stacked landmarks [X_nys; X_obs])"""
        return self.mode in ("predictive_covariance", "weighted_predictive_covariance", "gspace")

    @property
    def stationary(self):
        return self.family != _lib.TANIMOTO


def _cls(obj):
    return type(obj).__name__


def _describe_covar(covar):
    """-> (family, inv_ls, outputscale) or None."""
    outputscale = 1.0
    base = covar
    if _cls(covar) == "ScaleKernel":
        os_ = covar.outputscale
        if torch.is_tensor(os_):
            if os_.numel() != 1:
                return None          # batched kernels: generic path
            os_ = float(os_.detach().reshape(-1)[0])
        outputscale = float(os_)
        base = covar.base_kernel
    if getattr(base, "active_dims", None) is not None:
        return None
    name = _cls(base)
    if name == "TanimotoKernel":
        return _lib.TANIMOTO, None, outputscale
    if name in ("RBFKernel", "MaternKernel"):
        ls = base.lengthscale
        if not torch.is_tensor(ls) or ls.dim() > 2 or (ls.dim() == 2 and ls.shape[0] != 1):
            return None
        inv_ls = (1.0 / ls.detach().to(torch.float64)).reshape(-1)
        if name == "RBFKernel":
            return _lib.RBF, inv_ls, outputscale
        nu = float(base.nu)
        fam = {0.5: _lib.MATERN12, 1.5: _lib.MATERN32, 2.5: _lib.MATERN52}.get(nu)
        if fam is None:
            return None
        return fam, inv_ls, outputscale
    return None


def _covariance_cache(model):
    """``get_cov_cache`` of SOBER/_gp.py:255-278: W = S S^T with S the prediction strategy's covar cache."""
    x_obs = model.train_inputs[0]
    try:
        root = model.prediction_strategy.covar_cache
    except Exception:
        model.eval()
        model(x_obs[0].unsqueeze(0))
        root = model.prediction_strategy.covar_cache
    root = root.detach()
    return root @ root.T, x_obs.detach()


def _introspect_gspace(kernel) -> Optional[KernelSpec]:
    """``ScaleMmltGP.gspace_kernel`` (SOBER/BASQ/_scale_mmlt.py:256-275), the bound method BASQ.quadrature hands to
    ``recombination`` (SOBER/BASQ/_basq.py:55-67):  mu_g(x) mu_g(y) (exp(cov_h(x, y)) - 1)  with the h-space GP
    ``owner.model``.  Non-linear in the covariance: served by the POST mode of K1 (csrc/group_accumulate.cu)."""
    owner = getattr(kernel, "__self__", None)
    if owner is None or getattr(kernel, "__name__", "") != "gspace_kernel" or not hasattr(owner, "model"):
        return None
    from ._predict import describe_gp
    gp = describe_gp(owner.model)
    if gp is None:
        return None
    spec = gp.kernel
    spec.mode = "gspace"
    spec.x_obs, spec.woodbury, spec.alpha = gp.x_obs, gp.woodbury, gp.alpha
    spec.mean_const, spec.noise = gp.mean_const, gp.noise
    try:
        spec.jitter = float(torch.as_tensor(getattr(owner, "jitter", 0.0)).reshape(-1)[0])
    except Exception:
        return None
    return spec


def introspect(kernel) -> Optional[KernelSpec]:
    gs = _introspect_gspace(kernel)
    if gs is not None:
        return gs
    model = getattr(kernel, "model", None)
    mode = getattr(kernel, "mode", None)
    if model is None or mode not in ("kernel", "predictive_covariance", "weighted_predictive_covariance"):
        return None
    covar = getattr(model, "covar_module", None)
    if covar is None:
        return None
    desc = _describe_covar(covar)
    if desc is None:
        return None
    family, inv_ls, outputscale = desc
    if not math.isfinite(outputscale):
        return None
    d = None
    if inv_ls is not None and inv_ls.numel() > 1:
        d = int(inv_ls.numel())
    spec = KernelSpec(family, d, inv_ls, outputscale, mode)
    if spec.posterior:
        try:
            spec.woodbury, spec.x_obs = _covariance_cache(model)
        except Exception:
            return None
    if spec.weighted:
        # m(x) cov(x, y) m(y), SOBER/_kernel.py:33-47: linear in the covariance, so the candidates' weights carry m(x_i)
        # and the basis carries m(z_l); needs the mean cache and a constant mean (predict_mean, SOBER/_gp.py:240-253)
        try:
            spec.alpha = model.prediction_strategy.mean_cache.detach().reshape(-1)
            const = getattr(model.mean_module, "constant", None)
            if const is None or spec.alpha.numel() != spec.x_obs.shape[0]:
                return None
            spec.mean_const = float(torch.as_tensor(const).detach().reshape(-1)[0])
        except Exception:
            return None
    return spec
