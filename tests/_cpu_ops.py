"""TEST DOUBLE for ``sober_b200._ops.CudaOps``: the same interface computed with plain torch on the CPU.

Lives under tests/ on purpose -- the product has no CPU path.  It lets the CPU suite exercise the HOST logic of
``sober_b200._rchq.Recombiner`` (grouping, remainder handling, closed-form compaction, sharding over gloo)
against the oracle / golden fixtures without a GPU.  The arithmetic mirrors the CUDA kernels' formulas
(expanded distance on prepared points), not the oracle's, so it also cross-checks the data layouts.
"""
import math

import torch

from oracle import rchq as oracle
from sober_b200._ops import LandmarkTable, PointSet  # plain containers, no CUDA needed

RBF, MATERN12, MATERN32, MATERN52, TANIMOTO, TANIMOTO_BITS, HAMMING_LUT = range(7)


def kernel_values(dot, xn, zn, family):
    """dot: (m, L) = x . zt ; xn: (m, 1) ; zn: (1, L).  Same convention as csrc/common.cuh: the family constant
    is already folded into the coordinates (sober_b200._lib.FAMILY_SCALE)."""
    if family in (TANIMOTO, TANIMOTO_BITS):
        return ((dot + 1e-6) / (1e-6 + xn + zn - dot)).clamp_min(0)
    d2 = (xn + zn + dot).clamp_min(0)
    if family == RBF:
        return torch.exp(-d2)
    r = d2.clamp_min(1e-30).sqrt()
    if family == MATERN12:
        return torch.exp(-r)
    if family == MATERN32:
        return (1 + r) * torch.exp(-r)
    return (1 + r + d2 / 3.0) * torch.exp(-r)


class TorchOps:
    def __init__(self):
        self.device = torch.device("cpu")
        self.variant = 0

    def f64(self, t):
        return t.to(device=self.device, dtype=torch.float64).contiguous()

    def prepare_points(self, X, center, inv_ls):
        n, d = X.shape
        ldp = (d + 2) // 2 * 2
        P = torch.zeros((n, ldp), dtype=torch.float64)
        P[:, :d] = (X - center) * inv_ls
        P[:, d] = (P[:, :d] ** 2).sum(-1)
        return PointSet(P, ldp, P[:, d], ldp, n, d)

    def make_records(self, X, center, inv_ls, idx=None, mu=None):
        d = X.shape[1]
        rows = X if idx is None else X[idx.long()]
        m = rows.shape[0]
        ldr = (d + 3) // 2 * 2
        rec = torch.zeros((m, ldr), dtype=torch.float64)
        rec[:, :d] = (rows - center) * inv_ls
        rec[:, d] = (rec[:, :d] ** 2).sum(-1)
        rec[:, d + 1] = 1.0 if mu is None else mu
        return PointSet(None, 0, None, 0, m, d, rec=rec, ldr=ldr)

    def raw_points(self, X):
        n, d = X.shape
        return PointSet(X, X.stride(0), (X * X).sum(-1), 1, n, d)

    def pack_bits(self, X):
        """Test double: the 'words' are the float rows themselves (the popcount path is exercised on the GPU)."""
        ok = bool(((X == 0) | (X == 1)).all())
        return X, (X != 0).sum(-1).to(torch.float64), ok

    def compact_nonzero(self, mu):
        idx = torch.nonzero(mu != 0).reshape(-1).to(torch.int32)
        return idx, mu[idx.long()].clone(), int(idx.numel())

    def group_accumulate(self, pts, lm, idx, mu, n_local, pos0, ES, S, n_global=None, rec=None, unit_weights=False,
                         post=None):
        L, d = lm.L, lm.d
        at = torch.zeros((S, L), dtype=torch.float64)
        totw = torch.zeros(S, dtype=torch.float64)
        if n_local == 0:
            return at, totw
        if rec is not None:
            x, xn = rec[:n_local, :d], rec[:n_local, d].reshape(-1, 1)
            w = torch.ones(n_local, dtype=torch.float64) if unit_weights else rec[:n_local, d + 1]
        else:
            rows = torch.arange(n_local) if idx is None else idx[:n_local].long()
            w = torch.ones(n_local, dtype=torch.float64) if mu is None else mu[:n_local]
            x = pts.rows[rows, :d]
            xn = pts.xn[rows].reshape(-1, 1)
        if lm.family == HAMMING_LUT:
            ham = (xn + lm.zn.reshape(1, -1) - 2.0 * (x @ lm.zt.T)).round().long()
            kv = lm.lut[ham] * w.reshape(-1, 1)
        elif post is not None:
            # non-linear posterior mode (csrc/group_accumulate.cu, POST): expm1(s k(z, x) - <aw_l, kx_i>)
            kx, aw = post
            base = kernel_values(x @ lm.zt.T, xn, lm.zn.reshape(1, -1), lm.family) * lm.outputscale
            kv = torch.expm1(base - kx[rows] @ aw.T) * w.reshape(-1, 1)
            pos = pos0 + torch.arange(n_local)
            at.index_add_(0, pos % S, kv)
            inside = pos < ES
            totw.index_add_(0, (pos % S)[inside], w[inside])
            return at, totw
        else:
            kv = kernel_values(x @ lm.zt.T, xn, lm.zn.reshape(1, -1), lm.family) * w.reshape(-1, 1)
        pos = pos0 + torch.arange(n_local)
        at.index_add_(0, pos % S, kv)
        inside = pos < ES
        totw.index_add_(0, (pos % S)[inside], w[inside])
        return at * lm.outputscale, totw

    def group_accumulate_gram(self, G, mu, pos_begin, ES, S, At, totw):
        m = G.shape[1]
        w = torch.ones(m, dtype=torch.float64) if mu is None else mu
        pos = pos_begin + torch.arange(m)
        At.index_add_(0, pos % S, (G * w.reshape(1, -1)).T.contiguous())
        if totw is not None:
            inside = pos < ES
            totw.index_add_(0, (pos % S)[inside], w[inside])

    def car_eliminate(self, basis_rows, mass, want_pivots=False, exact=True):
        removed = oracle.eliminate(basis_rows.T.clone(), mass, oracle.Factory())
        if want_pivots:
            k = basis_rows.shape[0]
            piv = torch.full((max(k, 1),), -1, dtype=torch.int32)
            piv[:len(removed)] = torch.tensor(removed, dtype=torch.int32)
            return piv, torch.tensor([len(removed)], dtype=torch.int32)

    def update_compact(self, idx, mu, n_local, pos0, ES, S, wstar, totw, rank, K, tail_keep, new_pos0, n_out,
                       rec=None, d=0):
        pos = pos0 + torch.arange(n_local)
        g = torch.where(pos < ES, pos % S, torch.full_like(pos, S - 1))
        keep = torch.where(pos < ES, wstar[g] > 0, torch.full_like(pos, bool(tail_keep), dtype=torch.bool))
        dst = torch.where(pos < ES, (pos // S) * K + rank[g].long(), (ES // S) * K + (pos - ES)) - new_pos0
        idx_out = torch.zeros(n_out, dtype=torch.int32)
        mu_out = torch.zeros(n_out, dtype=torch.float64)
        idx_out[dst[keep]] = idx[:n_local][keep]
        mu_out[dst[keep]] = (mu[:n_local][keep] * wstar[g[keep]]) / totw[g[keep]]
        rec_out = None
        if rec is not None:
            rec_out = torch.zeros((n_out, rec.shape[1]), dtype=torch.float64)
            rec_out[dst[keep]] = rec[:n_local][keep]
            rec_out[:, d + 1] = mu_out
        return idx_out, mu_out, rec_out

    def kmeans_assign(self, X, centroids, chunk=4096):
        out = torch.empty(X.shape[0], dtype=torch.int64)
        for s in range(0, X.shape[0], chunk):
            d2 = ((X[s:s + chunk, None, :] - centroids[None, :, :]) ** 2).sum(-1)
            out[s:s + chunk] = d2.argmin(1)
        return out

    def scatter_result(self, dst, idx, w):
        dst.zero_()
        dst[idx] = w
