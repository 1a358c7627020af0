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
    """feats (S x n), mass (S,) -> weights (S,) with zeros at eliminated points; preserves [1 feats]^T mass."""
    ones = torch.ones((feats.shape[0], 1), dtype=feats.dtype, device=feats.device)
    design = torch.cat([ones, feats], dim=1)
    rows = nullspace(design) if nullspace is not None else nullspace_rows(design, how)
    out = mass.clone().contiguous()
    if rows.shape[0] > 0:
        ops.car_eliminate(rows, out)
    return out
