"""Nystrom basis of the landmark block (SOBER/_rchq.py:34-39): top-q approximate eigenvectors of the gated Gram
through the randomised range finder of ``torch.svd_lowrank`` (Halko et al. alg. 5.1, niter=2, no oversampling).

``U = -(Q @ svd(Q^T K).U)^T`` -- NOT scaled by the singular values, exactly like the reference.
"""
import torch

# Tests can set this to inject a test matrix drawn elsewhere (e.g. on the CPU generator the oracle used).
_injected_test_matrix = None


def draw_test_matrix(rows, cols, dtype, device):
    """The single random draw of the path: ``torch.randn(L, q)`` on the device's generator, as the reference's
    ``torch.svd_lowrank`` call does (consumes the global RNG stream in the same way)."""
    if _injected_test_matrix is not None:
        return _injected_test_matrix.to(device=device, dtype=dtype)
    return torch.randn(rows, cols, dtype=dtype, device=device)


def _orthonormal_basis(y, how):
    """Q of the thin QR of y (m x q, m >= q).

    ``householder``: torch.linalg.qr (cuSOLVER geqrf + orgqr: ~1.2 ms for 1000 x 199 on B200, an unblocked panel
    kernel).  ``cholqr2``: Cholesky-QR applied twice -- two Gram GEMMs, two q x q Cholesky factorisations, two
    triangular solves; mathematically the same Q up to column signs (the QR factorisation of a full-rank matrix is
    unique up to signs, and the final basis U = Q svd(Q^T K).U does not see them), numerically orthonormal to
    rounding while cond(y) < ~1e8.  Falls back to Householder when a Cholesky factorisation breaks down
    (numerically rank-deficient y, e.g. a low-dimensional RBF Gram)."""
    if how == "cholqr2" and y.shape[0] >= y.shape[1]:
        q = y
        for _ in range(2):
            gram = q.mH @ q
            chol, info = torch.linalg.cholesky_ex(gram)
            if int(info) != 0:
                return torch.linalg.qr(y).Q
            q = torch.linalg.solve_triangular(chol.mH, q, upper=True, left=False)
        return q
    return torch.linalg.qr(y).Q


def lowrank_basis(gram, rank, niter=2, qr="householder"):
    """-(Q @ svd(Q^T K).U)^T of torch.svd_lowrank, with the same single random draw.

    ``qr="householder"`` and no injected test matrix: ``torch.svd_lowrank`` itself (parity mode).  Otherwise the same
    sequence spelled out, with the orthonormalisations by ``_orthonormal_basis`` and the final SVD taken on the
    triangular factor of (Q^T K)^T (left singular vectors of B = right singular vectors of R when B^T = Q_B R):
    q x q instead of q x L Jacobi rotations."""
    if _injected_test_matrix is None and qr == "householder":
        left, _, _ = torch.svd_lowrank(gram, q=rank, niter=niter)
        return -1 * left.T
    size = gram.shape[-1]
    probe = draw_test_matrix(size, rank, gram.dtype, gram.device)
    q = _orthonormal_basis(gram @ probe, qr)
    for _ in range(niter):
        q = _orthonormal_basis(gram.mH @ q, qr)
        q = _orthonormal_basis(gram @ q, qr)
    small = q.mH @ gram                                   # (q x L)
    if qr == "cholqr2" and small.shape[0] <= small.shape[1]:
        gs = small @ small.mH
        chol, info = torch.linalg.cholesky_ex(gs)         # B^T = Q_B R with R = chol^H: only R is needed
        if int(info) == 0:
            # second pass for a numerically clean R:  B^T = Q1 R1, Q1 = Q2 R2  ->  R = R2 R1
            q1t = torch.linalg.solve_triangular(chol, small, upper=False)          # Q1^T  (q x L)
            chol2, info2 = torch.linalg.cholesky_ex(q1t @ q1t.mH)
            if int(info2) == 0:
                r_full = chol2.mH @ chol.mH                                          # upper triangular R
                _, _, vh = torch.linalg.svd(r_full)
                return -1 * (q @ vh.mH).T
    u_small, _, _ = torch.linalg.svd(small, full_matrices=False)
    return -1 * (q @ u_small).T
