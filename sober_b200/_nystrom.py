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


def lowrank_basis(gram, rank, niter=2):
    if _injected_test_matrix is None:
        left, _, _ = torch.svd_lowrank(gram, q=rank, niter=niter)
        return -1 * left.T
    size = gram.shape[-1]
    probe = draw_test_matrix(size, rank, gram.dtype, gram.device)
    q = torch.linalg.qr(gram @ probe).Q
    for _ in range(niter):
        q = torch.linalg.qr(gram.mH @ q).Q
        q = torch.linalg.qr(gram @ q).Q
    small = q.mH @ gram
    u_small, _, _ = torch.linalg.svd(small, full_matrices=False)
    return -1 * (q @ u_small).T
