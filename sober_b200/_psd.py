"""The PSD gate applied to the Nystrom block (semantics of SOBER/_utils.py:117-157), written for the device.

``gate="reference"``: Cholesky succeeds AND the matrix is bitwise symmetric AND every eigenvalue returned by the
general (non-symmetric) eigensolver has a non-negative real part -- the reference's test, evaluated with torch on
the current device.  ``gate="cholesky"``: Cholesky only (for a symmetric matrix the other two clauses are
redundant up to borderline rounding; the non-symmetric ``eig`` of a 2000 x 2000 block costs seconds).
"""
import warnings

import torch


def passes(mat, gate):
    try:
        _, info = torch.linalg.cholesky_ex(mat)
        if int(info) != 0:
            return False
        if gate == "cholesky":
            return True
        if not bool((mat == mat.T).all()):
            return False
        return bool((torch.linalg.eig(mat)[0].real >= 0).all())
    except Exception:      # the reference treats ANY failure of the test as "not PSD" (SOBER/_utils.py:128-129)
        return False


def repair(cov, gate, assume_asymmetric=False, max_iter=10):
    """Returns the matrix the Nystrom factorisation should see.

    ``assume_asymmetric``: treat the input as having failed the first test (what happens to every gpytorch-built
    Gram of practical size, whose matmul-based distances are asymmetric in the last bit -- SURVEY.md TL;DR 8);
    used when the Gram comes from the CUDA kernel, which is symmetric by construction.
    """
    if not assume_asymmetric and passes(cov, gate):
        return cov
    warnings.warn("Estimated covariance matrix was not positive semi-definite. Conveting...")
    cov = torch.nan_to_num(cov)
    cov = torch.sqrt(cov * cov.T)
    if passes(cov, gate):
        return cov
    size = cov.size(0)
    bump = torch.full((size,), 1e-5, dtype=cov.dtype, device=cov.device)
    rounds = 0
    diag = torch.arange(size, device=cov.device)
    while not passes(cov, gate):
        cov[diag, diag] += bump
        bump = bump * 2
        rounds += 1
        if rounds > max_iter:
            return cov.diag().diag()
    return cov
