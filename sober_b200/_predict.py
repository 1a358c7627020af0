"""GP posterior over the candidate set and the LFI acquisition measure -- SURVEY.md 8(f) row 1, the step immediately
before the recombination over the same N candidates:

  ``predict(test_x, model)``   SOBER/_gp.py:212-238   mean_i = c + k(x_i, X_obs) alpha,
                                                      var_i  = k(x_i, x_i) - k(x_i, X_obs) W k(X_obs, x_i) + noise
  ``PI.lfi(X_cand)``           SOBER/_pi.py:20-38     pi_i   = Phi((mean_i - eta) / sqrt(var_i))
  (importance weights ``pi(X) / prior.pdf(X)``: SOBER/_sampler.py:173-187, 351-382 -- the prior's own pdf stays where it is)

The reference hands the N x n_obs cross-covariance to gpytorch, which materialises it (and, under ``fast_pred_var``,
replaces W by a LOVE low-rank approximation).  Here the candidates stream through in chunks:

  1. K = k(X_chunk, X_obs): ONE K1 launch (``csrc/group_accumulate.cu`` in Gram mode, the same fused distance +
     nonlinearity code the recombination passes use; X_obs are the "landmarks");
  2. T = K W: the one dense contraction of the row, n_obs^2 FMAs per candidate -- a plain library DGEMM (cuBLAS);
  3. ``sober_gp_rows`` (``csrc/predict.cu``): both dot products, variance clamp and the normal CDF in one streaming pass.

W is the exact (K_obs + noise I)^-1 the prediction strategy caches (``covar_cache`` root, as ``get_cov_cache`` reads it,
SOBER/_gp.py:255-278): exact variances, not LOVE's approximation.  Kernel arithmetic of gpytorch itself is "parity
unpinned" (library absent from this image).
"""
import torch

from . import _lib
from ._kernel_spec import KernelSpec, _describe_covar

_MIN_VARIANCE = 1e-10      # gpytorch.settings.min_variance for float64 (MultivariateNormal.variance clamps at it)


class GPSpec:
    def __init__(self, kernel, x_obs, alpha, woodbury, mean_const, noise):
        self.kernel, self.x_obs, self.alpha, self.woodbury = kernel, x_obs, alpha, woodbury
        self.mean_const, self.noise = mean_const, noise


def describe_gp(model):
    """Duck-typed read of an exact-GP model (gpytorch ``ExactGP`` / botorch ``SingleTaskGP`` or the test stand-in):
    covariance module, training inputs, the prediction strategy's caches, constant mean, homoskedastic noise.
    ``None`` when something is not of that form (the caller keeps the reference's own code path then)."""
    try:
        desc = _describe_covar(model.covar_module)
        if desc is None:
            return None
        family, inv_ls, outputscale = desc
        x_obs = model.train_inputs[0].detach()
        try:
            strategy = model.prediction_strategy
            if strategy is None:
                raise AttributeError
        except Exception:
            model.eval()
            model(x_obs[0].unsqueeze(0))
            strategy = model.prediction_strategy
        root = strategy.covar_cache.detach()
        alpha = strategy.mean_cache.detach().reshape(-1)
        const = getattr(model.mean_module, "constant", None)
        if const is None:
            return None
        const = float(torch.as_tensor(const).detach().reshape(-1)[0])
        noise = torch.as_tensor(model.likelihood.noise).detach().reshape(-1)
        if noise.numel() != 1 or x_obs.dim() != 2 or alpha.numel() != x_obs.shape[0]:
            return None
        d = int(inv_ls.numel()) if (inv_ls is not None and inv_ls.numel() > 1) else None
        spec = KernelSpec(family, d, inv_ls, outputscale, "kernel")
        return GPSpec(spec, x_obs, alpha, root @ root.T, const, float(noise[0]))
    except Exception:
        return None


def gp_posterior(model, X, eta=None, ops=None, chunk=1 << 17, spec=None):
    """-> (mean, var, pi) float64 on the device; ``pi`` is None unless ``eta`` is given."""
    from ._rchq import Recombiner, _ops
    ops = ops or _ops()
    spec = spec or describe_gp(model)
    if spec is None:
        raise _lib.SoberB200Error("sober_b200: this model cannot be described as an exact GP with a constant mean and a "
                                  "RBF / Matern / Tanimoto kernel (see sober_b200/_predict.py)")
    dev = ops.device
    X = ops.f64(X)
    if X.dim() != 2 or X.shape[1] != spec.x_obs.shape[1]:
        raise ValueError("X (N, d) must have the dimension of the training inputs")
    ks = spec.kernel
    x_obs = ops.f64(spec.x_obs)
    d = X.shape[1]
    if ks.stationary:
        center = x_obs.mean(0).contiguous()
        inv_ls = (ops.f64(ks.inv_ls) * _lib.FAMILY_SCALE[ks.family]).expand(d).contiguous()
    else:
        center = torch.zeros(d, dtype=torch.float64, device=dev)
        inv_ls = torch.ones(d, dtype=torch.float64, device=dev)
    helper = Recombiner(ops)                       # landmark table / candidate layouts exactly as the recombination builds them
    table = helper._table(x_obs, ks, center, inv_ls)
    alpha = ops.f64(spec.alpha)
    w = ops.f64(spec.woodbury)
    n = X.shape[0]
    mean = torch.empty(n, dtype=torch.float64, device=dev)
    var = torch.empty(n, dtype=torch.float64, device=dev)
    pi = torch.empty(n, dtype=torch.float64, device=dev) if eta is not None else None
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        k_rows = helper._gram_T(helper._points(X[s:e], ks, center, inv_ls), table)        # (m x n_obs), scaled
        t_rows = k_rows @ w                                                               # library DGEMM
        ops.gp_rows(k_rows, t_rows, alpha, spec.mean_const, ks.outputscale, spec.noise, _MIN_VARIANCE,
                    0.0 if eta is None else float(eta), mean[s:e], var[s:e], None if pi is None else pi[s:e])
    return mean, var, pi


def predict(test_x, model):
    """Drop-in for ``SOBER._gp.predict`` (SOBER/_gp.py:212-238): ``(pred.mean, pred.variance)``."""
    mean, var, _ = gp_posterior(model, test_x)
    return mean, var


def current_maximum(model, spec=None):
    """``eta`` of SOBER/_pi.py:15: the largest posterior mean over the training inputs."""
    spec = spec or describe_gp(model)
    mean, _, _ = gp_posterior(model, spec.x_obs, spec=spec)
    return float(mean.max())


def pi_lfi(model, X_cand, eta=None, log=False):
    """Drop-in for ``PI.lfi`` (SOBER/_pi.py:20-38)."""
    spec = describe_gp(model)
    if eta is None:
        eta = current_maximum(model, spec)
    _, _, pi = gp_posterior(model, X_cand, eta=eta, spec=spec)
    if log:
        return (pi + torch.finfo().eps).log()      # torch.finfo() of the DEFAULT dtype, as the reference writes it
    return pi
