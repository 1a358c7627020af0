"""Oracle restatement (TEST INFRASTRUCTURE) of the GP posterior and the LFI acquisition measure, SURVEY.md 8(f) row 1:

  ``predict``  SOBER/_gp.py:212-238: ``model.likelihood(model(test_x))`` -> (mean, variance).  The arithmetic lives in
               gpytorch's exact prediction strategy (un-vendored, absent here: "parity unpinned"); restated from its
               documented behaviour: mean = c + k(x, X) mean_cache, covariance = k(x, x) - k(x, X) (S S^T) k(X, x) with
               S the ``covar_cache`` root, the Gaussian likelihood adds the noise, ``.variance`` clamps at
               ``settings.min_variance`` (1e-10 in float64).  Exact caches, not the LOVE approximation of fast_pred_var.
  ``lfi``      SOBER/_pi.py:20-38 -- pinned against the reference's own PI class (tests/golden/make_golden_pi.py runs the
               unmodified SOBER/_pi.py with ``predict`` bound to the function below).
"""
import torch

MIN_VARIANCE = 1e-10


def predict(test_x, model):
    x_obs = model.train_inputs[0]
    root = model.prediction_strategy.covar_cache
    k_xo = model.covar_module.forward(test_x, x_obs)
    mean = model.mean_module.constant + k_xo @ model.prediction_strategy.mean_cache
    if test_x.dim() == 2 and len(test_x) > 4096:        # the diagonal only: chunked so that nothing N x N is formed
        prior_var = torch.cat([torch.diagonal(model.covar_module.forward(test_x[s:s + 4096], test_x[s:s + 4096]))
                               for s in range(0, len(test_x), 4096)])
    else:                                                # (batch, S, d) inputs: one diagonal per batch entry
        prior_var = torch.diagonal(model.covar_module.forward(test_x, test_x), dim1=-2, dim2=-1)
    var = prior_var - ((k_xo @ root) ** 2).sum(-1) + model.likelihood.noise
    return mean, var.clamp_min(MIN_VARIANCE)


def lfi(x_cand, model, eta, log=False):
    """SOBER/_pi.py:20-38.  ``log=True`` states the evident intent of :35-36; the reference itself raises NameError there
    (the module never imports ``torch``)."""
    import torch.distributions as D
    mu_pred, var_pred = predict(x_cand, model)
    val = D.Normal(0, 1).cdf((mu_pred - eta) / var_pred.sqrt())
    if log:
        return (val + torch.finfo().eps).log()
    return val


def current_maximum(model):
    """``eta`` of SOBER/_pi.py:15."""
    mean, _ = predict(model.train_inputs[0], model)
    return mean.max().item()
