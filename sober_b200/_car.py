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


def projector_rows(design, thin_q=None, with_defect=False):
    """Null-space basis from the orthogonal projector: the trailing k columns of P = I - Q1 Q1^T, Q1 an orthonormal
    basis of range(design).  Spans null(design^T) whenever the leading n' x n' block of Q1 is non-singular (generic);
    NOT orthonormal -- the elimination only needs a basis, its ratio tests are invariant to column scaling.

    Everything is GEMM-shaped (no n'-step Householder sequence).  Columns of the design are normalised (does not
    change the null space); ONE Cholesky-QR pass gives Qt with Qt^T Qt = I + Delta, |Delta| ~ eps cond^2; the exact
    projector I - Qt (I + Delta)^-1 Qt^T is then formed with the Neumann series (I + Delta)^-1 = I - Delta + Delta^2
    (error |Delta|^3) -- three small GEMMs instead of a second Cholesky + triangular solve.  ``with_defect`` also
    returns |Delta|_F (a device scalar) so that the caller can fall back when it is not small.  The projector does not
    see how Q1 was orthonormalised: a Householder Q1 gives the same matrix (what the oracle-side check uses)."""
    from ._linalg import cholesky_upper, solve_right_upper
    pts, dim = design.shape
    eye_k = torch.eye(pts - dim, dtype=design.dtype, device=design.device)
    if thin_q is not None:
        rows = -(thin_q[dim:, :] @ thin_q.mH)
        rows[:, dim:] += eye_k
        return rows.contiguous()
    scaled = design / design.norm(dim=0, keepdim=True).clamp_min(1e-300)
    r, _ = cholesky_upper(scaled.mH @ scaled)                       # no host sync: NaNs surface in the caller's check
    qt = solve_right_upper(r, scaled)
    delta = qt.mH @ qt
    delta.diagonal().sub_(1.0)
    inv = delta @ delta - delta
    inv.diagonal().add_(1.0)                                         # I - Delta + Delta^2
    rows = -((qt[dim:, :] @ inv) @ qt.mH)                            # (k x S)
    rows[:, dim:] += eye_k
    rows = rows.contiguous()
    if with_defect:
        return rows, torch.linalg.matrix_norm(delta)
    return rows


def caratheodory(ops, feats, mass, how, nullspace=None, design=None):
    """feats (S x n), mass (S,) -> weights (S,) with zeros at eliminated points; preserves [1 feats]^T mass.
    ``design`` = [1 | feats] may be passed instead of ``feats`` (the projection kernel emits it directly).

    Small problems (the whole state fits the distributed shared memory of one thread-block cluster) run as ONE
    kernel, ``sober_car_cluster``: QR + null space + elimination for ``how="qr"``, the elimination alone -- with the
    reference's exact arithmetic -- on a torch-SVD / injected basis otherwise.  Larger ones take the null space from
    torch.linalg and the persistent multi-CTA elimination kernel."""
    if design is None:
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
    defect = None
    if nullspace is not None:
        rows = nullspace(design)
    elif how == "projector":
        rows, defect = projector_rows(design, with_defect=True)
    else:
        rows = nullspace_rows(design, how)
    exact = nullspace is not None or how == "svd"          # parity / injected bases keep the reference's rounding
    cols_fit = getattr(ops, "car_cols_fits", None)
    panel_fit = getattr(ops, "car_panel_fits", None)
    from ._settings import options
    if (not exact and panel_fit is not None and options.car_kernel != "legacy" and panel_fit(pts, rows.shape[0])):
        ops.car_panel(rows, out, nb_hint=options.car_panel_nb)   # panelled row-distributed cluster kernel
    elif cols_fit is not None and cols_fit(pts, rows.shape[0]):
        ops.car_cols(rows, out, exact=exact)                # column-distributed cluster kernel (fastest)
    elif fits is not None and fits(pts, dim, True):
        ops.car_cluster(out, basis_rows=rows, exact=exact)
    else:
        ops.car_eliminate(rows, out, exact=exact)
    if defect is not None:
        # poison the result when the one-pass Cholesky-QR was not accurate enough for the Neumann correction
        # (|Delta|^3 must stay below rounding): the caller's finite-check then redoes the step with Householder QR
        out = torch.where(defect < 1e-5, out, torch.full_like(out, float("nan")))
    return out


def needs_retry(how, kept_count, dim, finite):
    """The projector basis can be rank-deficient (singular leading block of Q1), its Cholesky-QR can break down or be
    too inaccurate (see ``caratheodory``); all show up as more than n' survivors or non-finite weights.  The caller
    checks this with the sync it does anyway and redoes the step with the Householder basis."""
    return how == "projector" and (kept_count > dim or not finite)


# ---------------------------------------------------------------------------------------------------------
# One Caratheodory step as a CUDA graph
# ---------------------------------------------------------------------------------------------------------
# Every iteration of the recombination loop reduces an (S x n) feature matrix with the same S and n: ~35 launches
# (column scaling, Gram, Cholesky, triangular solve, the Neumann-corrected projector, the elimination kernel, the
# survivor flags and ranks), 0.2 ms of GPU time that the host needs 0.3 ms to enqueue.  Captured once per shape and
# replayed, the step costs the host one launch.
_graphs = {}


_consts = {}


def _step_constants(S, dim, device):
    """-I and I (dim x dim) and [0 | I_k] (k x S): inputs of the addmm calls below, built once per shape."""
    key = (S, dim, str(device))
    c = _consts.get(key)
    if c is None:
        eye = torch.eye(dim, dtype=torch.float64, device=device)
        e2t = torch.zeros((S - dim, S), dtype=torch.float64, device=device)
        e2t[:, dim:] = torch.eye(S - dim, dtype=torch.float64, device=device)
        c = _consts[key] = (-eye, eye, e2t)
    return c


def projector_rows_sharded(comm, scaled):
    """The projector null space of ``_reduce_step_fused`` with its S-dimension work split over the ranks of ``comm``.

    Every rank holds the same column-scaled design ``scaled`` (S x dim) -- the Caratheodory step is replicated, its
    inputs are bit-identical after the all-reduce of the group sums -- and at C5 sizes (S = 2002, dim = 1001) the null
    space is 1.6 ms of GEMMs and a triangular solve per call that eight GPUs would each repeat.  Here rank r takes a
    contiguous block of rows of ``scaled``: its part of the Gram matrix (all-reduce), its rows of Q = scaled R^-1
    (all-gather), its part of Q^T Q (all-reduce), and a block of the k = S - dim null-space rows (all-gather).  The
    Cholesky factorisation and the dim x dim Neumann term stay replicated.  Collectives hand every rank the same bits, so
    the replicated elimination that follows still takes the same pivots on every rank.
    Returns (rows (k x S), delta (dim x dim)) like the replicated code path (same formulas; sums in a different order)."""
    from ._linalg import cholesky_upper, solve_right_upper
    S, dim = scaled.shape
    k = S - dim
    W, r = comm.world, comm.rank
    dev, f64 = scaled.device, torch.float64
    step = -(-S // W)
    lo, hi = min(S, r * step), min(S, (r + 1) * step)
    part = scaled[lo:hi]
    gram = part.mH @ part
    comm.all_reduce(gram)
    rr, _ = cholesky_upper(gram)
    q_loc = torch.zeros((step, dim), dtype=f64, device=dev)
    if hi > lo:
        q_loc[:hi - lo] = solve_right_upper(rr, part)
    q_all = torch.empty((W * step, dim), dtype=f64, device=dev)
    comm.all_gather(q_all, q_loc)
    qt = q_all[:S]
    delta = q_loc.mH @ q_loc                              # the padding rows are zero
    comm.all_reduce(delta)
    delta.diagonal().sub_(1.0)                            # Q^T Q - I
    eye = torch.eye(dim, dtype=f64, device=dev)
    inv = torch.addmm(eye - delta, delta, delta)          # I - Delta + Delta^2
    kstep = -(-k // W)
    klo, khi = min(k, r * kstep), min(k, (r + 1) * kstep)
    mine = torch.zeros((kstep, S), dtype=f64, device=dev)
    if khi > klo:
        blk = mine[:khi - klo]
        blk[:, dim + klo:dim + khi] = torch.eye(khi - klo, dtype=f64, device=dev)
        blk.addmm_(qt[dim + klo:dim + khi] @ inv, qt.mH, alpha=-1.0)
    rows_all = torch.empty((W * kstep, S), dtype=f64, device=dev)
    comm.all_gather(rows_all, mine)
    return rows_all[:k], delta


def _reduce_step_fused(ops, feats, mass, divide, comm=None):
    """The fast-mode step with its bookkeeping folded into two hand-written kernels and the GEMM epilogues:
    ``car_prepare`` (barycentres + ones column + column scaling), Gram / Cholesky / triangular solve, the Neumann-
    corrected projector as three addmm calls (the -I, +I and [0 | I] terms ride on the GEMMs' beta operand), the
    panelled elimination, ``car_summary`` (accuracy check of the projector, survivor counts and ranks).
    14 launches instead of ~35; same arithmetic as ``projector_rows`` + ``_reduce_step``.
    With ``comm``: the null space is computed by ``projector_rows_sharded`` (multi-GPU, large S)."""
    from ._linalg import cholesky_upper, solve_right_upper
    from ._settings import options
    S, n = feats.shape
    dim = n + 1
    scaled = ops.car_prepare(feats, mass if divide else None)           # (S x dim), unit columns
    if comm is not None:
        rows, delta = projector_rows_sharded(comm, scaled)
    else:
        neg_eye, eye, e2t = _step_constants(S, dim, feats.device)
        r, _ = cholesky_upper(scaled.mH @ scaled)
        qt = solve_right_upper(r, scaled)
        delta = torch.addmm(neg_eye, qt.mH, qt)                             # Q^T Q - I
        inv = torch.addmm(eye - delta, delta, delta)                        # I - Delta + Delta^2
        rows = torch.addmm(e2t, qt[dim:, :] @ inv, qt.mH, alpha=-1.0)       # trailing k columns of I - Q (I + Delta)^-1 Q^T
    out = mass.clone()
    ops.car_panel(rows, out, nb_hint=options.car_panel_nb)
    summary, rank = ops.car_summary(out, delta)
    return out, None, summary, rank


def _reduce_step(ops, feats, mass, divide=False):
    """feats (S x n), mass (S,) -> weights (S,), kept mask, host summary, exclusive rank of the kept.
    ``divide``: feats are group SUMS, the barycentres are feats / mass (SOBER/_rchq.py:166).
    The host summary is int32 [inclusive cumulative count of kept groups (S) | all weights finite (1)]: one small
    device-to-host copy gives the host everything it needs (``KeepMap.from_summary``)."""
    from ._settings import options
    S, n = feats.shape
    pfits = getattr(ops, "car_panel_fits", None)
    if (hasattr(ops, "car_prepare") and options.car_kernel != "legacy" and pfits is not None and S > n + 1
            and S <= 4096 and pfits(S, S - n - 1)):
        return _reduce_step_fused(ops, feats, mass, divide)
    if divide:
        feats = feats / mass.unsqueeze(1)
    wfull = caratheodory(ops, feats, mass, "projector")
    kept = wfull > 0
    k32 = kept.to(torch.int32)
    cum = torch.cumsum(k32, 0).to(torch.int32)
    rank = cum - k32
    summary = torch.cat([cum, torch.isfinite(wfull).all().reshape(1).to(torch.int32)])
    return wfull, kept, summary, rank


class _ReduceGraph:
    def __init__(self, ops, feats, mass, divide=False):
        self.feats = torch.empty_like(feats)
        self.mass = torch.empty_like(mass)
        self.feats.copy_(feats)
        self.mass.copy_(mass)
        saved, ops.timing = ops.timing, None            # event records cannot be timed inside a capture
        try:
            torch.cuda.synchronize(feats.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.out = _reduce_step(ops, self.feats, self.mass, divide)
        finally:
            ops.timing = saved
        self.launches = 0

    def __call__(self, feats, mass):
        self.feats.copy_(feats)
        self.mass.copy_(mass)
        self.graph.replay()
        return self.out


def reduce_step(ops, feats, mass, use_graph=True, divide=False, comm=None):
    """``_reduce_step``, replayed from a CUDA graph from the second call with the same shapes on (the outputs are then
    static buffers, valid until the next call).  Falls back to eager execution for good if the capture fails.
    ``comm`` (a multi-rank communicator) with S >= options.car_shard_min: the step runs eagerly with the projector null
    space split over the ranks (``projector_rows_sharded``; collectives are kept out of graph capture)."""
    fits = getattr(ops, "car_cols_fits", None)
    S, n = feats.shape
    pfits = getattr(ops, "car_panel_fits", None)
    from ._settings import options
    if (comm is not None and getattr(comm, "world", 1) > 1 and feats.is_cuda and S >= options.car_shard_min
            and hasattr(ops, "car_prepare") and options.car_kernel != "legacy" and pfits is not None and S > n + 1
            and S <= 4096 and pfits(S, S - n - 1)):
        t0 = ops._begin("car_step_graph")                # same stage label as the graph replays (bench.py's stage table)
        out = _reduce_step_fused(ops, feats, mass, divide, comm)
        ops._end("car_step_graph", t0, S - n - 1)
        return out
    if (not use_graph or not feats.is_cuda or fits is None or S <= n + 1
            or not (fits(S, S - n - 1) or (pfits is not None and pfits(S, S - n - 1)))):
        return _reduce_step(ops, feats, mass, divide)
    key = (feats.device, S, n, divide)
    entry = _graphs.get(key)
    if entry is None:
        _graphs[key] = "seen"
        return _reduce_step(ops, feats, mass, divide)
    from . import _linalg
    if entry == "seen":
        before, before_la = ops.launches, _linalg.launches
        try:
            entry = _graphs[key] = _ReduceGraph(ops, feats, mass, divide)
            # kernels of ours inside the graph (bench.py's launch count): replays run them without passing the wrappers
            entry.launches = (ops.launches - before, _linalg.launches - before_la)
        except Exception as err:                        # capture unsupported here: stay eager, but say so once
            import warnings
            warnings.warn("sober_b200: CUDA-graph capture of the Caratheodory step failed (%s); running it eagerly"
                          % (str(err).splitlines()[0] if str(err) else type(err).__name__))
            _graphs[key] = "eager"
            ops.launches, _linalg.launches = before, before_la
            return _reduce_step(ops, feats, mass, divide)
        ops.launches, _linalg.launches = before, before_la      # the capture itself executed nothing
    if entry == "eager":
        return _reduce_step(ops, feats, mass, divide)
    ops.launches += entry.launches[0]
    _linalg.launches += entry.launches[1]
    t0 = ops._begin("car_step_graph")                   # the whole replay: null space + elimination + survivor ranks
    out = entry(feats, mass)
    ops._end("car_step_graph", t0, S - n - 1)
    return out
