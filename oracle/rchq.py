"""Oracle restatement of the RCHQ recombination path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, stage by stage, with the same torch operations in the same order so that a CPU run is
bit-identical to the reference (pinned by tests/golden, see tests/test_oracle_golden.py):

  ``recombination``            SOBER/_rchq.py:5-31, 42-48
  ``psd_gate`` / ``repair_psd`` SOBER/_utils.py:117-129 / 131-157
  ``nystrom_basis``            SOBER/_rchq.py:34-39
  ``reduce_measure``           SOBER/_rchq.py:51-221   (Mod_Tchernychova_Lyons)
  ``caratheodory``             SOBER/_rchq.py:224-270  (Tchernychova_Lyons_CAR)

Every stage reports its intermediates to an optional ``trace`` callback ``trace(stage_name, dict)``; the
golden generator and the stage-wise parity tests hang off that.  ``nullspace`` lets a test swap the
null-space basis (default: trailing rows of ``Vh`` from a full ``torch.linalg.svd`` exactly like
SOBER/_rchq.py:231-234) so the product's *fast* basis can be fed through the reference's elimination rule.
"""
import warnings

import torch


class Factory:
    """The handful of ``TensorManager`` constructors the path uses (SOBER/_utils.py:37-60)."""

    def __init__(self, device=None, dtype=torch.float64):
        self.device = torch.device("cpu") if device is None else torch.device(device)
        self.dtype = dtype

    def ones(self, *shape):
        return torch.ones(*shape).to(self.device, self.dtype)

    def zeros(self, *shape):
        return torch.zeros(*shape).to(self.device, self.dtype)

    def arange(self, n):
        return torch.arange(n).to(self.device)


# ------------------------------------------------------------------------------------------------
# PSD gate   (SOBER/_utils.py:117-157)
# ------------------------------------------------------------------------------------------------
def psd_gate(mat):
    """True iff Cholesky succeeds AND the matrix is bitwise symmetric AND every eigenvalue of the general
    (non-symmetric) eigen-solver has non-negative real part.  SOBER/_utils.py:125-129."""
    try:
        torch.linalg.cholesky(mat)
        symmetric = (mat == mat.T).all()
        return bool(symmetric and (torch.linalg.eig(mat)[0].real >= 0).all())
    except Exception:
        return False


def repair_psd(cov, fac, max_iter=10, trace=None):
    """SOBER/_utils.py:131-157: geometric-mean symmetrisation then escalating diagonal jitter
    (1e-5, 2e-5, 4e-5, ... cumulative), diagonal-only fallback after ``max_iter`` + 1 additions."""
    rounds = -1
    if not psd_gate(cov):
        warnings.warn("Estimated covariance matrix was not positive semi-definite. Conveting...")
        cov = torch.nan_to_num(cov)
        cov = torch.sqrt(cov * cov.T)
        rounds = 0
        if not psd_gate(cov):
            m = cov.size(0)
            bump = fac.ones(m) * 1e-5
            while not psd_gate(cov):
                cov[range(m), range(m)] += bump
                bump *= 2
                rounds += 1
                if rounds > max_iter:
                    cov = cov.diag().diag()
                    break
    if trace is not None:
        trace("psd", {"rounds": rounds, "K": cov})
    return cov


# ------------------------------------------------------------------------------------------------
# Nystrom basis   (SOBER/_rchq.py:34-39)
# ------------------------------------------------------------------------------------------------
def nystrom_basis(landmarks, rank, kernel, fac, trace=None):
    gram = kernel(landmarks, landmarks)
    if trace is not None:
        trace("gram", {"K_raw": gram.clone()})
    gram = repair_psd(gram, fac, trace=trace)
    left, sing, _ = torch.svd_lowrank(gram, q=rank)
    basis = -1 * left.T
    if trace is not None:
        trace("basis", {"U": basis, "S": sing})
    return basis


# ------------------------------------------------------------------------------------------------
# Caratheodory elimination   (SOBER/_rchq.py:224-270)
# ------------------------------------------------------------------------------------------------
def svd_nullspace(design):
    """Trailing rows of Vh of the FULL svd of design^T  (SOBER/_rchq.py:231-234).  Returns Phi (N x (N-n))."""
    pts, dim = design.shape
    _, _, vh = torch.linalg.svd(design.T)
    return vh[-(pts - dim):, :].T


def eliminate(phi, mass, fac):
    """The pivoting loop SOBER/_rchq.py:237-266 on a given null-space basis ``phi`` (N x k); ``mass`` is
    updated in place.  Returns the list of eliminated positions."""
    removed = []
    for _ in range(phi.shape[1]):
        lead = phi[:, 0]
        pos = lead > 0
        if pos.sum() == 0:                      # guard added upstream on 7 Aug 2023 (:241-242)
            break
        ratio = fac.zeros(len(mass))
        ratio[pos] = mass[pos] / lead[pos]
        cand = fac.arange(len(mass))[pos]
        pivot = cand[torch.argmin(ratio[pos])]
        removed.append(int(pivot))
        mass[:] = mass - ratio[pivot] * lead
        mass[pivot] = 0.0
        rest = phi[:, 1:]
        rest = rest - torch.matmul(rest[pivot].unsqueeze(1), lead.unsqueeze(1).T).T / lead[pivot]
        rest[pivot, :] = 0.0
        phi = rest
    return removed


def caratheodory(feats, mass, fac, nullspace=None, trace=None):
    """Reduce ``N`` weighted points with features ``feats`` (N x n) to at most n+1, preserving
    ``[1 feats]^T mass``.  ``mass`` is consumed (callers pass a clone, SOBER/_rchq.py:85,174)."""
    design = torch.cat([fac.ones(feats.size(0)).unsqueeze(0).T, feats], dim=1)
    phi = svd_nullspace(design) if nullspace is None else nullspace(design)
    if trace is not None:
        trace("car_in", {"X": feats, "mu": mass.clone(), "Phi": phi.clone()})
    eliminate(phi, mass, fac)
    keep = mass > 0
    w, idx = mass[keep], fac.arange(design.shape[0])[keep]
    if trace is not None:
        trace("car_out", {"w": w, "idx": idx})
    return w, idx


def _objective_step(feat_rows, obj_vals, w, idx, fac):
    """The extra direction taken when ``calc_obj`` is given (SOBER/_rchq.py:87-106 and 177-196):
    one more null-space move on the n+2 surviving points, signed to increase the objective."""
    pts = torch.cat((feat_rows, fac.ones(1, len(idx))), 0)
    _, _, vh = torch.linalg.svd(pts)
    direction = vh[-1]
    if torch.dot(obj_vals, direction) < 0:
        direction = -direction
    pos = direction > 0
    ratio = fac.zeros(len(w))
    ratio[pos] = w[pos] / direction[pos]
    cand = fac.arange(len(w))[pos]
    pivot = cand[torch.argmin(ratio[pos])]
    w = w - ratio[pivot] * direction
    w[pivot] = 0.0
    live = w > 0
    return w[live], idx[live]


# ------------------------------------------------------------------------------------------------
# the measure-reduction loop   (SOBER/_rchq.py:51-221)
# ------------------------------------------------------------------------------------------------
def group_moments(cands, basis, landmarks, kernel, mass, alive, groups, fac, obj=None, trace=None):
    """One grouped pass (SOBER/_rchq.py:116-166): returns barycentres (S x n[+1]), group masses (S,),
    the (E, S) index table and the tail indices."""
    rows = int(len(alive) / groups)
    covered = groups * rows
    table = alive[:covered].reshape(rows, groups)
    gram = kernel(landmarks, cands[table]) * mass[table].unsqueeze(1)            # (E, L, S)
    col_sums = fac.zeros(basis.shape[1], groups)
    col_sums += gram.sum(axis=0)
    tail = alive[covered:]
    if len(tail) > 0:
        # the remainder is first folded into the leading group columns (:128-136) ...
        extra = kernel(landmarks, cands[tail]) * mass[tail].unsqueeze(0)
        col_sums += torch.cat((extra, fac.zeros(basis.shape[1], groups - len(tail))), dim=1)
    if obj is not None:
        obj_sums = fac.zeros(1, groups)
        obj_sums += (obj[table].unsqueeze(1) * mass[table].unsqueeze(1)).sum(axis=0)
        if len(tail) > 0:
            o_extra = obj[tail].unsqueeze(0) * mass[tail].unsqueeze(0)
            obj_sums += torch.cat((o_extra, fac.zeros(1, groups - len(tail))), dim=1)
    proj = basis @ col_sums
    if obj is not None:
        proj = torch.cat((proj, obj_sums), 0)
    bary = proj.T
    totals = torch.sum(mass[table], 0)
    if len(tail):
        # ... and then counted once more into the last group (:153-164)
        tail_feats = basis @ kernel(landmarks, cands[tail])
        if obj is not None:
            tail_feats = torch.cat((tail_feats, torch.reshape(obj[tail], (1, -1))), 0)
        bary[-1] += torch.multiply(tail_feats.T, mass[tail].unsqueeze(1)).sum(axis=0)
        totals[-1] += torch.sum(mass[tail], 0)
    if trace is not None:
        trace("group", {"A": col_sums.clone(), "totw": totals.clone(), "Xt_unnormalised": bary.clone(),
                        "R": len(alive), "E": rows})
    bary = torch.divide(bary, totals.unsqueeze(0).T)
    return bary, totals, table, tail


def reduce_measure(cands, basis, landmarks, kernel, fac, mass=None, calc_obj=None, nullspace=None, trace=None):
    total_pts = len(cands)
    n, _ = basis.shape
    groups = 2 * (n + 1)
    if mass is None:
        mass = fac.ones(total_pts) / total_pts
    alive = fac.arange(total_pts)
    alive = alive[mass != 0]
    left = len(alive)
    obj = None if calc_obj is None else -1 * calc_obj(cands)

    while True:
        if left <= n + 1:
            idx = fac.arange(len(mass))[mass > 0]
            return mass[idx], idx

        if n + 1 < left <= groups:
            feats = basis @ kernel(landmarks, cands[alive])
            if obj is not None:
                feats = torch.cat((feats, torch.reshape(obj[alive], (1, -1))), 0)
                feats_raw = torch.clone(feats[:-1])
            w, idx = caratheodory(feats.T, torch.clone(mass[alive]), fac, nullspace, trace)
            if obj is not None:
                w, idx = _objective_step(feats_raw[:, idx], obj[idx], w, idx, fac)
            alive = alive[idx]
            mass[:] = 0.0
            mass[alive] = w
            return mass[mass > 0], alive

        bary, totals, table, tail = group_moments(
            cands, basis, landmarks, kernel, mass, alive, groups, fac, obj, trace)
        if obj is not None:
            bary_raw = torch.clone(bary[:, :n])
            obj_bary = bary[:, -1:].reshape(-1)
        w, kept = caratheodory(bary, torch.clone(totals), fac, nullspace, trace)
        if obj is not None:
            w, kept = _objective_step(bary_raw[kept].T, obj_bary[kept], w, kept, fac)

        survivors = table[:, kept].reshape(-1)
        drop = fac.ones(table.shape[1]).to(torch.bool)
        drop[kept] = 0
        mass[table[:, drop].reshape(-1)] = 0.0
        scaled = torch.multiply(mass[table[:, kept]], w)
        scaled = torch.divide(scaled, totals[kept])
        mass[survivors] = scaled.reshape(-1)

        last = fac.arange(len(kept))[(kept == groups - 1) != 0]
        if len(last) > 0:
            # the tail rides with the last barycentre (:208-215)
            scaled = torch.multiply(mass[tail], w[last])
            scaled = torch.divide(scaled, totals[kept[last]])
            mass[tail] = scaled
            survivors = torch.cat([survivors, tail])
        else:
            mass[tail] = 0.0
        alive = torch.clone(survivors)
        left = len(alive)
        if trace is not None:
            trace("update", {"alive": alive.clone(), "mass_alive": mass[alive].clone()})


def recombination(pts_rec, pts_nys, num_pts, kernel, device=None, dtype=None, init_weights=None,
                  calc_obj=None, nullspace=None, trace=None):
    """Same signature and return contract as SOBER/_rchq.py:5-31 -> (idx, w); ``init_weights`` is mutated in
    place into the sparse solution just like the reference (:109-110, :203-218).  ``device`` / ``dtype`` pick
    the working factory here (the reference takes them from SOBER._settings instead, :30)."""
    fac = Factory(pts_rec.device if device is None else device, torch.float64 if dtype is None else dtype)
    basis = nystrom_basis(pts_nys, num_pts - 1, kernel, fac, trace)
    w, idx = reduce_measure(pts_rec, basis, pts_nys, kernel, fac, mass=init_weights, calc_obj=calc_obj,
                            nullspace=nullspace, trace=trace)
    return idx, w


# ------------------------------------------------------------------------------------------------
# quality metric used by parity tests (the reference never computes it)
# ------------------------------------------------------------------------------------------------
def mmd_squared(kernel, cands, mass, idx, w, chunk=4096):
    """``w^T K_bb w - 2 w^T K_bN mu + mu^T K_NN mu`` with the kernel callable, evaluated in chunks."""
    batch = cands[idx]
    t1 = w @ kernel(batch, batch) @ w
    t2 = cands.new_zeros(())
    t3 = cands.new_zeros(())
    for s in range(0, len(cands), chunk):
        blk, m = cands[s:s + chunk], mass[s:s + chunk]
        t2 = t2 + w @ kernel(batch, blk) @ m
        for s2 in range(0, len(cands), chunk):
            t3 = t3 + m @ kernel(blk, cands[s2:s2 + chunk]) @ mass[s2:s2 + chunk]
    return t1 - 2 * t2 + t3


def projector_nullspace(design):
    """Restatement, with LAPACK on the CPU, of the null-space basis the product's *fast* mode uses
    (sober_b200/_car.py::projector_rows): the trailing k columns of the orthogonal projector I - Q1 Q1^T, Q1 an
    orthonormal basis of range(design) (columns normalised first; the projector does not depend on how Q1 was
    orthonormalised).  Pass as ``recombination(..., nullspace=projector_nullspace)``: the reference's algorithm
    (SOBER/_rchq.py:224-270) with this basis in place of the arbitrary one its full SVD returns (:231-234)."""
    pts, dim = design.shape
    q1 = torch.linalg.qr(design / design.norm(dim=0, keepdim=True)).Q
    phi = -(q1 @ q1[dim:, :].T)
    phi[dim:, :] += torch.eye(pts - dim, dtype=design.dtype, device=design.device)
    return phi

