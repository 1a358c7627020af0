"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the weighted-KDE density of the reference
(``SOBER/_wkde.py:109-145``, ``WeightedKernelDensityEstimation.pdf``), the SURVEY.md §8(f) row "weighted KDE prior
update / pdf".  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.

Pinned against the reference itself: ``tests/golden/make_golden_wkde.py`` runs the unmodified class under a stub
``SOBER`` package and stores (Xobs, weights, covariance, bounds, queries, pdf) in ``tests/golden/wkde_*.npz``;
``tests/test_wkde.py`` asserts this restatement reproduces those values bit for bit.

    pdf(x) = sum_j w_j N(x - c_j; 0, Sigma)              (``safe_mvn_prob``: MultivariateNormal.log_prob(.).exp(),
                                                           SOBER/_utils.py:171-194)
    rows of out-of-bound queries are zeroed when bounds are given (:131-136); with ``compute_cdf`` the weights are
    divided by the per-centre truncation constants (:138-139).
"""
import torch
from torch.distributions.multivariate_normal import MultivariateNormal


def pdf(centres, weights, covariance, queries, bounds=None, constant=None, chunk=500_000):
    """centres (n_kde, d), weights (n_kde,), covariance (d, d), queries (N, d) -> (N,).  Same op sequence as the
    reference, including its (n_X * n_kde, d) difference tensor, evaluated in chunks of queries."""
    n_kde, d = centres.shape
    mvn = MultivariateNormal(torch.zeros(d, dtype=centres.dtype), covariance)      # SOBER/_utils.py:159-169
    out = []
    step = max(1, chunk // max(n_kde, 1))
    w = weights if constant is None else weights / constant
    for s in range(0, len(queries), step):
        q = queries[s:s + step]
        diff = (centres.repeat(len(q), 1, 1) - q.unsqueeze(1)).reshape(n_kde * len(q), d)     # :121-123
        dens = mvn.log_prob(diff).exp().reshape(len(q), n_kde)                                 # :125-129
        if bounds is not None:
            dens[(q < bounds[0]).any(axis=1)] = 0.0                                            # :133-136
            dens[(q > bounds[1]).any(axis=1)] = 0.0
        out.append(w @ dens.T)                                                                 # :138-143
    return torch.cat(out) if out else torch.zeros(0, dtype=centres.dtype)
