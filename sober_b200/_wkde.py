"""Weighted-KDE density on the B200 path -- SURVEY.md §8(f) row 3, ``WeightedKernelDensityEstimation.pdf``
(``SOBER/_wkde.py:109-145``): ``pdf(x_i) = sum_j w_j N(x_i - c_j; 0, Sigma)`` over N queries and n_kde centres, rows of
out-of-bound queries zeroed.

The reference materialises the (N * n_kde, d) difference tensor and calls ``MultivariateNormal.log_prob`` on it (hence its
5e5-row splitting, ``SOBER/_utils.py:182-191``).  Here it is the recombination's K1 kernel with the roles swapped:

* whitening ``u = L^-1 x`` (``Sigma = L L^T``, the same Cholesky factor ``MultivariateNormal`` takes) turns the Gaussian
  into ``exp(-|u_i - v_j|^2 / 2)``: the RBF family of ``csrc/group_accumulate.cu`` with unit lengthscale;
* the n_kde CENTRES are the weighted "candidates" (record layout, weight w_j), dealt round-robin into S = 4 groups
  so that the kernel's 4-group register tile is full, and the N QUERIES are the "landmarks": one launch fills the
  (4 x N) accumulator, whose column sums times ``(2 pi)^(-d/2) / prod(diag L)`` are the densities.

Nothing of size N * n_kde is ever stored; no new CUDA code (the S = 1 / small-S shapes are what the remainder pass of
every recombination iteration already exercises).
"""
import math

import torch

from . import _lib
from ._ops import LandmarkTable

_GROUPS = 4            # = REC_TG of csrc/group_accumulate.cu: a full register tile
# queries per launch: grid.y of K1 is limited to 65535 blocks of 512 "landmarks" (record kernel, d <= 8) or of 64 (tiled
# kernel of the indexed layout, d > 8)
_MAX_QUERIES_RECORDS = 1 << 24
_MAX_QUERIES_TILED = 65535 * 64 // 2


def _max_queries(d):
    return _MAX_QUERIES_RECORDS if d <= _lib.RECORD_MAX_D else _MAX_QUERIES_TILED


def wkde_pdf(centres, weights, covariance, queries, bounds=None, constant=None, ops=None):
    """centres (n_kde, d), weights (n_kde,), covariance (d, d), queries (N, d) -> densities (N,) float64 on the device.
    ``bounds`` (2, d): queries outside get 0 (SOBER/_wkde.py:131-136); ``constant`` (n_kde,): per-centre truncation
    constants dividing the weights (``compute_cdf=True``, :138-139).
    ``covariance`` must be symmetric positive definite, as the estimator's is after the gate it passes at construction
    (``_compute_covariance``, :98-107); the second, idempotent pass of that gate inside ``safe_mvn_register``
    (SOBER/_utils.py:159-169) is therefore not repeated here."""
    if ops is None:
        from ._rchq import _ops
        ops = _ops()
    c, w, q = ops.f64(centres), ops.f64(weights), ops.f64(queries)
    cov = ops.f64(covariance)
    if c.dim() != 2 or q.dim() != 2 or c.shape[1] != q.shape[1] or w.shape != (c.shape[0],) or cov.shape != (c.shape[1],) * 2:
        raise ValueError("centres (n_kde, d), weights (n_kde,), covariance (d, d), queries (N, d) expected")
    n_kde, d = c.shape
    if constant is not None:
        w = w / ops.f64(constant)
    out = torch.zeros(q.shape[0], dtype=torch.float64, device=q.device)
    if n_kde == 0 or q.shape[0] == 0:
        return out
    chol = torch.linalg.cholesky(cov)                                   # what MultivariateNormal(mu, cov) factors
    inv_t = torch.linalg.solve_triangular(chol, torch.eye(d, dtype=torch.float64, device=cov.device), upper=False).T
    scale = math.exp(-0.5 * d * math.log(2.0 * math.pi)) / torch.diagonal(chol).prod()
    cw = c @ inv_t                                                      # rows L^-1 c_j
    pad = (-n_kde) % _GROUPS                                            # zero-weight copies: no remainder group
    if pad:
        cw = torch.cat([cw, cw[:1].expand(pad, d)], 0)
        w = torch.cat([w, torch.zeros(pad, dtype=torch.float64, device=w.device)])
    n_pad = n_kde + pad
    center = cw.mean(0).contiguous()
    inv_ls = torch.full((d,), _lib.FAMILY_SCALE[_lib.RBF], dtype=torch.float64, device=cw.device)   # exp(-|.|^2 / 2)
    cw, w = cw.contiguous(), w.contiguous()
    if d <= _lib.RECORD_MAX_D:
        pts, rec, mu = None, ops.make_records(cw, center, inv_ls, None, w).rec, None
    else:
        pts, rec, mu = ops.prepare_points(cw, center, inv_ls), None, w
    step = _max_queries(d)
    for s in range(0, q.shape[0], step):
        v = ((q[s:s + step] @ inv_t) - center) * inv_ls
        table = LandmarkTable((-2.0 * v).contiguous(), (v * v).sum(-1).contiguous(), _lib.RBF, 1.0)
        at, _ = ops.group_accumulate(pts, table, None, mu, n_pad, 0, n_pad, _GROUPS, rec=rec)
        out[s:s + step] = at.sum(0) * scale
    if bounds is not None:
        b = ops.f64(bounds)
        out[(q < b[0]).any(1) | (q > b[1]).any(1)] = 0.0
    return out


def pdf_of(kde, queries, ops=None):
    """``kde.pdf(queries)`` for a reference ``WeightedKernelDensityEstimation`` (duck-typed: ``Xobs``, ``weights``,
    ``covariance``, ``bounds``, ``compute_cdf``, ``constant``)."""
    constant = getattr(kde, "constant", None) if (getattr(kde, "compute_cdf", False) and kde.bounds is not None) else None
    return wkde_pdf(kde.Xobs, kde.weights, kde.covariance, queries, bounds=kde.bounds, constant=constant, ops=ops)
