"""Caratheodory reduction (SOBER/_rchq.py:224-270): null-space basis on the host side of the C ABI (torch.linalg
on the device), the k dependent elimination steps in one persistent CUDA kernel (csrc/car_eliminate.cu)."""
import torch


def nullspace_rows(design, how):
    """Rows spanning null(design^T) as a contiguous (k x S) matrix; row c plays column c of the reference's Phi.

    ``svd``: trailing rows of Vh of the full SVD of design^T -- SOBER/_rchq.py:231-234 verbatim.
    ``qr`` : trailing columns of the complete Householder Q of design (orthonormal basis of the same space;
             reproducible across LAPACK/cuSOLVER up to rounding, unlike the SVD's arbitrary basis).
    """
    pts, dim = design.shape
    if how == "svd":
        _, _, vh = torch.linalg.svd(design.T)
        return vh[-(pts - dim):, :].contiguous()
    if how == "qr":
        q = torch.linalg.qr(design, mode="complete").Q
        return q[:, dim:].T.contiguous()
    if how == "projector":
        return projector_rows(design)
    raise ValueError(how)


def projector_rows(design, thin_q=None):
    """Null-space basis from the orthogonal projector: the trailing k columns of P = I - Q1 Q1^T, Q1 an orthonormal
    basis of range(design).  Spans null(design^T) whenever the leading n' x n' block of Q1 is non-singular (generic);
    NOT orthonormal -- the elimination only needs a basis, its ratio tests are invariant to column scaling.

    Everything is GEMM-shaped (no n'-step Householder sequence): columns of the design are normalised (does not
    change the null space), Q1 comes from Cholesky-QR applied twice (the projector does not see column signs, so a
    Householder Q1 gives the same matrix -- that is what the oracle-side check uses), then one GEMM."""
    from ._nystrom import _orthonormal_basis
    pts, dim = design.shape
    if thin_q is None:
        scaled = design / design.norm(dim=0, keepdim=True).clamp_min(1e-300)
        thin_q = _orthonormal_basis(scaled, "cholqr2", check=False)     # no host sync; see caratheodory()
    rows = -(thin_q[dim:, :] @ thin_q.mH)                       # (k x S): - Q1[n':, :] Q1^T
    rows[:, dim:] += torch.eye(pts - dim, dtype=design.dtype, device=design.device)
    return rows.contiguous()


def caratheodory(ops, feats, mass, how, nullspace=None):
    """feats (S x n), mass (S,) -> weights (S,) with zeros at eliminated points; preserves [1 feats]^T mass.

    Small problems (the whole state fits the distributed shared memory of one thread-block cluster) run as ONE
    kernel, ``sober_car_cluster``: QR + null space + elimination for ``how="qr"``, the elimination alone -- with the
    reference's exact arithmetic -- on a torch-SVD / injected basis otherwise.  Larger ones take the null space from
    torch.linalg and the persistent multi-CTA elimination kernel."""
    ones = torch.ones((feats.shape[0], 1), dtype=feats.dtype, device=feats.device)
    design = torch.cat([ones, feats], dim=1).contiguous()
    pts, dim = design.shape
    out = mass.clone().contiguous()
    if pts <= dim:
        return out
    fits = getattr(ops, "car_cluster_fits", None)
    if nullspace is None and how == "qr" and fits is not None and fits(pts, dim, False):
        ops.car_cluster(out, design=design)
        return out
    rows = nullspace(design) if nullspace is not None else nullspace_rows(design, how)
    exact = nullspace is not None or how == "svd"          # parity / injected bases keep the reference's rounding
    cols_fit = getattr(ops, "car_cols_fits", None)
    if cols_fit is not None and cols_fit(pts, rows.shape[0]):
        ops.car_cols(rows, out, exact=exact)                # column-distributed cluster kernel (fastest)
    elif fits is not None and fits(pts, dim, True):
        ops.car_cluster(out, basis_rows=rows, exact=exact)
    else:
        ops.car_eliminate(rows, out, exact=exact)
    return out


def needs_retry(how, kept_count, dim, finite):
    """The projector basis can be rank-deficient (singular leading block of Q1) or its Cholesky-QR can break down;
    both show up as more than n' survivors or non-finite weights.  The caller checks this with the sync it does
    anyway and redoes the step with the Householder basis."""
    return how == "projector" and (kept_count > dim or not finite)
