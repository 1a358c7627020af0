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
    raise ValueError(how)


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
    exact = nullspace is not None or how != "qr"          # parity / injected bases keep the reference's rounding
    if fits is not None and fits(pts, dim, True):
        ops.car_cluster(out, basis_rows=rows, exact=exact)
    else:
        ops.car_eliminate(rows, out, exact=exact)
    return out
