"""Nystrom basis of the landmark block (SOBER/_rchq.py:34-39): top-q approximate eigenvectors of the gated Gram
through the randomised range finder of ``torch.svd_lowrank`` (Halko et al. alg. 5.1, niter=2, no oversampling).

``U = -(Q @ svd(Q^T K).U)^T`` -- NOT scaled by the singular values, exactly like the reference.
"""
import torch

from ._linalg import cholesky_upper, solve_right_upper

# Tests can set this to inject a test matrix drawn elsewhere (e.g. on the CPU generator the oracle used).
_injected_test_matrix = None


def draw_test_matrix(rows, cols, dtype, device):
    """The single random draw of the path: ``torch.randn(L, q)`` on the device's generator, as the reference's
    ``torch.svd_lowrank`` call does (consumes the global RNG stream in the same way)."""
    if _injected_test_matrix is not None:
        return _injected_test_matrix.to(device=device, dtype=dtype)
    return torch.randn(rows, cols, dtype=dtype, device=device)


class SideStream:
    """Runs independent bulk work beside a host-driven sequence of small kernels.

    ``with SideStream(launch, device, bulk_stream):`` -- on entry ``launch()`` is called with ``bulk_stream`` current
    (after making it wait for everything enqueued so far); it returns the tensors it produced.  The body then runs on
    the original stream; on exit that stream waits for the bulk work.  ``bulk_stream`` is confined to all SMs but a
    few (``sober_partition_stream``, a CUDA green context), so the body's kernels always find free SMs.  (Two ordinary
    streams do not overlap here, whatever their priorities: the bulk kernel's CTAs retire in waves ~0.4 ms apart and a
    small kernel only gets an SM then -- measured.)"""

    debug = None   # set to a list to collect {bulk_end, body_end} in ms after the fork (forces a sync)

    def __init__(self, launch, device, bulk_stream):
        self.launch, self.device, self.bulk = launch, device, bulk_stream

    def __enter__(self):
        self.main = torch.cuda.current_stream(self.device)
        timing = SideStream.debug is not None
        ready = torch.cuda.Event(enable_timing=timing)
        ready.record(self.main)
        self.bulk.wait_event(ready)
        with torch.cuda.stream(self.bulk):
            for t in self.launch():
                t.record_stream(self.main)          # allocated on the bulk stream, consumed on the main one
            if timing:
                self.ev = [ready, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
                self.ev[1].record(self.bulk)
        return self

    def __exit__(self, *exc):
        if SideStream.debug is not None:
            self.ev[2].record(self.main)
            torch.cuda.synchronize(self.device)
            SideStream.debug.append({"bulk_end": self.ev[0].elapsed_time(self.ev[1]),
                                     "body_end": self.ev[0].elapsed_time(self.ev[2])})
        self.main.wait_stream(self.bulk)
        return False


def _orthonormal_basis(y, how, check=True, passes=2):
    """Q of the thin QR of y (m x q, m >= q).

    ``householder``: torch.linalg.qr (cuSOLVER geqrf + orgqr: ~1.2 ms for 1000 x 199 on B200, an unblocked panel
    kernel).  ``check=False`` skips the (synchronising) breakdown test.  ``cholqr2``: Cholesky-QR applied twice -- two Gram GEMMs, two q x q Cholesky factorisations, two
    triangular solves; mathematically the same Q up to column signs (the QR factorisation of a full-rank matrix is
    unique up to signs, and the final basis U = Q svd(Q^T K).U does not see them), numerically orthonormal to
    rounding while cond(y) < ~1e8.  Falls back to Householder when a Cholesky factorisation breaks down
    (numerically rank-deficient y, e.g. a low-dimensional RBF Gram).  ``passes=1`` leaves Q orthonormal only to
    eps * cond(y)^2 -- enough for the intermediate bases of the power iteration, which exist for stability alone."""
    if how == "cholqr2" and y.shape[0] >= y.shape[1]:
        q = y
        for _ in range(passes):
            gram = q.mH @ q
            r, info = cholesky_upper(gram)
            # check=False: no host sync; a breakdown leaves NaNs that the caller detects downstream
            if check and int(info) != 0:
                return torch.linalg.qr(y).Q
            q = solve_right_upper(r, q)
        return q
    return torch.linalg.qr(y).Q


def lowrank_basis(gram, rank, niter=2, qr="householder", rotate=True, probe=None):
    """-(Q @ svd(Q^T K).U)^T of torch.svd_lowrank, with the same single random draw.

    ``qr="householder"`` and no injected test matrix: ``torch.svd_lowrank`` itself (parity mode).  Otherwise the same
    sequence spelled out, with the orthonormalisations by ``_orthonormal_basis`` and the final SVD taken on the
    triangular factor of (Q^T K)^T (left singular vectors of B = right singular vectors of R when B^T = Q_B R):
    q x q instead of q x L Jacobi rotations.

    ``rotate=False`` returns Q^T itself, without the q x q rotation by the left singular vectors of Q^T K: the rows span
    the same space (the rotation is orthogonal and complete -- no singular vector is dropped, q = rank), which is all
    that the recombination sees when its null spaces come from the orthogonal projector (``_car.projector_rows`` is a
    function of range([1 | features]) alone; the terminal branches and the objective step likewise).  The rotation is
    the most expensive N-independent item of the call (eigh + refinement: 2.7 ms of 20 at BASELINE configs[1])."""
    if _injected_test_matrix is None and qr == "householder" and probe is None:
        left, _, _ = torch.svd_lowrank(gram, q=rank, niter=niter)
        return -1 * left.T
    size = gram.shape[-1]
    if probe is None:                                     # (a caller that may need to redo the call draws it itself)
        probe = draw_test_matrix(size, rank, gram.dtype, gram.device)
    # range(Q) after the last orthonormalisation is range(K (K^T K)^niter Omega) whatever the intermediate bases
    # were: those only keep the columns from collapsing onto the dominant eigenvector, for which one Cholesky-QR
    # pass (orthonormal to eps cond^2) does as well as two.  The last basis is orthonormalised to rounding.
    last = 2 if rotate else 1                              # without the rotation only span(Q) matters
    q = _orthonormal_basis(gram @ probe, qr, passes=1 if niter > 0 else last)
    for it in range(niter):
        q = _orthonormal_basis(gram.mH @ q, qr, passes=1)
        q = _orthonormal_basis(gram @ q, qr, passes=last if it == niter - 1 else 1)
    if not rotate:
        return q.T.contiguous()
    small = q.mH @ gram                                   # B = Q^T K  (q x L)
    if qr == "cholqr2" and small.shape[0] <= small.shape[1]:
        u_small = _left_singular_vectors(small)
    else:
        u_small, _, _ = torch.linalg.svd(small, full_matrices=False)
    return -1 * (q @ u_small).T


def _left_singular_vectors(b, max_sweeps=12, tol=1e-13):
    """Left singular vectors of b (q x L, q <= L), descending, WITHOUT a Jacobi SVD (cuSOLVER gesvdj: 5 ms at
    199 x 1000 and 100-200 ms at 999 x 2000 on B200 -- the largest N-independent item of the call).

    1. eigh of the q x q Gram b b^T (1.9 / 12 ms): eigenvectors U0, accurate only to eps * cond(b)^2 / relgap.
    2. First-order corrections that work on the FACTOR, like one-sided Jacobi does: with c = U^T b (rows nearly
       orthogonal, norms ~ sigma_i) the entries g_ij = <c_i, c_j> carry an error eps * sigma_1 * sigma_j, so the
       rotation angles x_ij = g_ij / (g_jj - g_ii) are accurate to eps * cond(b) / relgap -- the accuracy of the
       Jacobi SVD itself.  U <- orth(U (I + X)): all pairs rotated at once with GEMMs, quadratically convergent once
       the angles are small; repeated until max |x_ij| < tol (2-4 sweeps for cond(b) up to ~1e6).
    Near-degenerate pairs (relative gap below 1e-9) are left alone: their individual vectors are ill-defined in the
    reference as well.  If the iteration does not settle (cond(b)^2 beyond what eigh can seed), the Jacobi SVD of
    torch.linalg.svd is used."""
    gram = b @ b.mH
    _, u = torch.linalg.eigh(gram)
    u = u.flip(-1)                                        # descending, like svd
    prev = float("inf")
    for sweep in range(max_sweeps):
        c = u.mH @ b
        g = c @ c.mH
        dg = torch.diagonal(g)
        gap = dg.unsqueeze(0) - dg.unsqueeze(1)           # gap[i, j] = g_jj - g_ii
        ok = gap.abs() > 1e-9 * (dg.unsqueeze(0) + dg.unsqueeze(1))
        x = torch.where(ok, g / torch.where(ok, gap, torch.ones_like(gap)), torch.zeros_like(g))
        x.diagonal().zero_()
        size = x.abs().max()
        x = x.clamp(-0.25, 0.25)
        u = u + u @ x                                     # U (I + X), X skew-symmetric to first order
        # re-orthonormalise: U^T U = I + D with D = O(|X|^2).  Small rotations: (I + D)^-1/2 = I - D/2 + 3 D^2/8 by
        # GEMMs alone (error |D|^3); large ones (first sweeps of a badly seeded problem): one Cholesky-QR pass
        status = size.to(torch.float64).reshape(1).tolist()          # the one host sync of the sweep
        if status[0] < 1e-3:
            d = u.mH @ u
            d.diagonal().sub_(1.0)
            u = u - 0.5 * (u @ d) + 0.375 * (u @ (d @ d))
        else:
            r, info = cholesky_upper(u.mH @ u)
            if int(info) != 0:
                break
            u = solve_right_upper(r, u)
        status = [0.0, status[0]]
        # converged, or stagnated at the noise floor eps * cond(b) / relgap of the closest pair (clustered spectrum):
        # the Jacobi SVD has the same intrinsic sensitivity there
        stalled = sweep >= 2 and status[1] < 1e-5 and status[1] > 0.25 * prev
        prev = status[1]
        if status[1] < tol or stalled:
            # order by the Rayleigh quotients actually reached (eigh's order can be off where it had no accuracy)
            order = torch.argsort(torch.diagonal((u.mH @ b) @ (u.mH @ b).mH), descending=True)
            return u[:, order]
    return torch.linalg.svd(b, full_matrices=False)[0]
