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
        return _all_eigenvalues_nonnegative(mat)
    except Exception:      # the reference treats ANY failure of the test as "not PSD" (SOBER/_utils.py:128-129)
        return False


def _all_eigenvalues_nonnegative(mat):
    """``(torch.linalg.eig(mat)[0].real >= 0).all()`` for a matrix that has just been found EXACTLY symmetric
    (``and`` short-circuits in SOBER/_utils.py:127: the general eigensolver only ever sees such a matrix).

    Both the symmetric and the general eigensolver are backward stable, and the eigenvalues of a symmetric (normal)
    matrix are perfectly conditioned: each computed eigenvalue is within c n eps |A| of an exact one.  So whenever the
    smallest eigenvalue from the symmetric solver clears that band by a wide margin (100 n eps |A|_F) the general solver
    returns the same sign pattern, and its verdict is known without running it (it is a hybrid CPU/GPU Hessenberg-QR
    iteration: the largest single cost of the reference's op sequence on a GPU, see profiles/r02_parity_breakdown.txt).
    Inside the band the verdict depends on rounding: the reference's own call decides."""
    size = mat.shape[-1]
    try:
        low = torch.linalg.eigvalsh(mat)[0]
        band = 100.0 * size * torch.finfo(mat.dtype).eps * torch.linalg.matrix_norm(mat)
        low, band = float(low), float(band)
        if low == low and band == band:                 # not NaN
            if low > band:
                return True
            if low < -band:
                return False
    except Exception:
        pass
    return bool((torch.linalg.eig(mat)[0].real >= 0).all())


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
    return escalate(cov, gate, max_iter)


def escalate(cov, gate, max_iter=10):
    """The jitter loop of the repair (SOBER/_utils.py:145-157), for a matrix that has just failed the test."""
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


# ---------------------------------------------------------------------------------------------------------
# Deferred test (fast mode): the L x L Cholesky of the gate only DECIDES; the range finder does not need its factor.
# It runs on a helper stream beside the range finder, which speculates on the usual outcome (test passed); the
# verdict is read when the basis is ready.  On failure the caller escalates and recomputes -- exactly what the
# blocking path would have produced.
# ---------------------------------------------------------------------------------------------------------
_helper_streams = {}


class PendingTest:
    def __init__(self, info, stream):
        self.info, self.stream = info, stream

    def passed(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.info.device).wait_stream(self.stream)
        return int(self.info) == 0


def repair_deferred(cov):
    """``repair(cov, "cholesky", assume_asymmetric=True)`` up to and including the launch of the first test:
    returns (matrix, PendingTest)."""
    warnings.warn("Estimated covariance matrix was not positive semi-definite. Conveting...")
    cov = torch.nan_to_num(cov)
    cov = torch.sqrt(cov * cov.T)
    if not cov.is_cuda:
        return cov, PendingTest(torch.linalg.cholesky_ex(cov)[1], None)
    dev = cov.device
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _helper_streams:
        _helper_streams[key] = torch.cuda.Stream(dev)
    helper, main = _helper_streams[key], torch.cuda.current_stream(dev)
    helper.wait_stream(main)
    with torch.cuda.stream(helper):
        info = torch.linalg.cholesky_ex(cov)[1]
    cov.record_stream(helper)
    info.record_stream(main)
    return cov, PendingTest(info, helper)
