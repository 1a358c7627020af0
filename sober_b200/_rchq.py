"""Drop-in for ``SOBER/_rchq.py``: same ``recombination(...)`` signature and ``(idx, w)`` contract
(SOBER/_rchq.py:5-31), B200-native execution.

Host logic (this file) mirrors the control flow of ``Mod_Tchernychova_Lyons`` (SOBER/_rchq.py:51-221); every
pass over the candidates is a hand-written sm_100a kernel reached through the C ABI (``_ops.CudaOps``):

  reference                                   here
  ------------------------------------------  --------------------------------------------------------------
  idx_story = arange(N)[mu != 0]     :63-65   ops.compact_nonzero            (stream compaction)
  kernel(pt_nys, samp[idx]) * mu, sum :124-136 ops.group_accumulate           (K1: never materialises (E,L,S))
  U_svd @ X_for_nys, / tot_weights   :148-166 one small GEMM  At @ Uext^T     (landmarks stacked for pred. cov.)
  Tchernychova_Lyons_CAR             :224-270 _car.caratheodory               (persistent elimination kernel)
  mu updates + idx_story rebuild     :198-221 ops.update_compact              (closed-form scatter, no scan)

The alive-list is kept COMPACT (row ids + weights of surviving points, ascending); "position" = rank in that
list, group = position mod S, exactly the reshape of SOBER/_rchq.py:118-123.  With ``torch.distributed``
enabled (``sober_b200.distributed``) candidates are row-sharded: each rank owns a contiguous range of positions
and the only collective per iteration is one all-reduce of the (S x L') accumulator.
"""
import itertools
import threading
import time

import contextlib

import torch

from . import _car, _lib, _nystrom, _psd
from ._kernel_spec import introspect
from ._ops import LandmarkTable, PointSet, record_stride
from ._settings import options


# ---------------------------------------------------------------------------------------------------------
# communication: nothing for one GPU, torch.distributed when sharded
# ---------------------------------------------------------------------------------------------------------
class SingleProcess:
    rank, world = 0, 1

    def all_reduce(self, tensor):
        return tensor

    def all_gather_ints(self, value, device):
        return [int(value)]

    def all_gather(self, out, part):
        out.copy_(part.reshape(out.shape))
        return out


class Sharded:
    """Row-sharded candidates over a torch.distributed group (NCCL on the GPU box, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_reduce(self, tensor):
        self.dist.all_reduce(tensor, group=self.group)
        return tensor

    def all_gather_ints(self, value, device):
        mine = torch.tensor([int(value)], dtype=torch.int64, device=device)
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        self.dist.all_gather_into_tensor(out, mine, group=self.group)
        return [int(v) for v in out.tolist()]

    def all_gather(self, out, part):
        """out (world * m, ...) <- the ranks' equally shaped parts (m, ...), in rank order."""
        self.dist.all_gather_into_tensor(out, part.contiguous(), group=self.group)
        return out


# ---------------------------------------------------------------------------------------------------------
# closed-form survivor counting (replaces the boolean-mask bookkeeping of SOBER/_rchq.py:198-221)
# ---------------------------------------------------------------------------------------------------------
class KeepMap:
    """Which positions survive an iteration, given the kept groups.  All host-side integers."""

    def __init__(self, kept_mask, S, ES):
        self.S, self.ES, self.E = S, ES, ES // S
        # cum[g] = number of kept groups below g
        self.cum = [0] + list(itertools.accumulate(1 if k else 0 for k in kept_mask[:S]))
        self.K = self.cum[S]
        self.tail_keep = bool(kept_mask[S - 1])

    @classmethod
    def from_summary(cls, summary, S, ES):
        """From the inclusive cumulative kept-count the device computed (``_car._reduce_step``): no Python loop over
        the groups (this runs between the host sync and the next kernel launch, with the GPU idle)."""
        self = cls.__new__(cls)
        self.S, self.ES, self.E = S, ES, ES // S
        self.cum = [0] + summary[:S]
        self.K = self.cum[S]
        self.tail_keep = self.cum[S] > self.cum[S - 1]
        return self

    def before(self, p):
        """Number of surviving positions strictly below global position p."""
        if p <= self.ES:
            return (p // self.S) * self.K + self.cum[p % self.S]
        return self.E * self.K + ((p - self.ES) if self.tail_keep else 0)


def _family_values(family, d2):
    """Kernel value as a function of the (family-scaled) squared distance -- the formulas of csrc/common.cuh."""
    if family == _lib.RBF:
        return torch.exp(-d2)
    r = d2.clamp_min(1e-30).sqrt()
    if family == _lib.MATERN12:
        return torch.exp(-r)
    if family == _lib.MATERN32:
        return (1 + r) * torch.exp(-r)
    return (1 + r + d2.clamp_min(1e-30) / 3.0) * torch.exp(-r)


class Alive:
    """This rank's part of the compact alive-list, in position order: row ids, weights and -- record layout --
    the record rows K1 streams."""

    def __init__(self, idx, mass, rec=None):
        self.idx, self.mass, self.rec = idx, mass, rec
        self.pending = None      # [(begin, end, event)]: record rows still to be built from candidates in flight to the device

    def tail(self, t0):
        return Alive(self.idx[t0:], self.mass[t0:], None if self.rec is None else self.rec[t0:])

    def part(self, a, b):
        return Alive(self.idx[a:b], self.mass[a:b], None if self.rec is None else self.rec[a:b])


class _StageClock:
    """Stage boundaries of one call.  ``options.stats = {}``: per-stage wall clock -- synchronises the device at every
    boundary, so it is a diagnostic, not something to leave on when timing the whole call.  ``options.nvtx`` (env
    SOBER_B200_NVTX=1): an NVTX range per stage instead ("sober_b200/k1", "sober_b200/car", ...; no synchronisation), for
    ``ncu --nvtx --nvtx-include`` / Nsight Systems timelines."""

    _NEXT = {"setup": "compact+records", "compact+records": "nystrom", "nystrom": "k1 | finish", "k1": "tail+project",
             "tail+project": "car", "car": "keepmap+update", "keepmap+update": "k1 | finish", "finish": None}

    def __init__(self, sink, device, nvtx=False):
        self.sink, self.device, self.t = sink, device, None
        self.nvtx = bool(nvtx) and device.type == "cuda"
        if sink is not None and device.type == "cuda":
            torch.cuda.synchronize(device)
            self.t = time.perf_counter()
        if self.nvtx:
            torch.cuda.nvtx.range_push("sober_b200/setup")

    def lap(self, name):
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
            nxt = self._NEXT.get(name)
            if nxt is not None:
                torch.cuda.nvtx.range_push("sober_b200/" + nxt)
            else:
                self.nvtx = False
        if self.t is None:
            return
        torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        self.sink[name] = self.sink.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now

    def close(self):
        if self.nvtx:
            torch.cuda.nvtx.range_pop()
            self.nvtx = False


class Recombiner:
    def __init__(self, ops, comm=None, opts=None, nullspace=None, basis=None, trace=None):
        self.ops = ops
        self.comm = comm or SingleProcess()
        self.opts = opts or options
        self.nullspace = nullspace        # test hook: design -> (k x S) rows
        self.basis = basis                # test hook: use this Nystrom basis U (n x L) instead of computing it
        self.trace = trace                # test hook: trace(stage, dict) with per-iteration intermediates
        self._bits = False                # set per call: "tanimoto" / "hamming" when {0,1} rows are bit-packed
        self._lut = None                  # Hamming path: kernel value per Hamming distance

    # -----------------------------------------------------------------------------------------------------
    # landmarks / Nystrom block
    # -----------------------------------------------------------------------------------------------------
    def _table(self, pts, spec, center, inv_ls):
        """LandmarkTable for raw landmark rows ``pts`` (L' x d)."""
        if self._bits:
            words, popc, ok = self.ops.pack_bits(pts)
            if not ok:
                raise ValueError("non-binary landmark rows on a bit-packed path")
            fam = _lib.HAMMING_LUT if self._bits == "hamming" else _lib.TANIMOTO_BITS
            return LandmarkTable(words, popc, fam, spec.outputscale, d=pts.shape[1], lut=self._lut)
        if spec.stationary:
            v = (pts - center) * inv_ls
            return LandmarkTable((-2.0 * v).contiguous(), (v * v).sum(-1).contiguous(), spec.family, spec.outputscale)
        return LandmarkTable(pts.contiguous(), (pts * pts).sum(-1).contiguous(), spec.family, spec.outputscale)

    def _use_records(self, spec, d):
        return (spec is not None and d <= _lib.RECORD_MAX_D and self.opts.k1_variant != 1
                and not getattr(spec, "gspace", False))          # the POST mode of K1 reads the indexed layout

    def _points(self, X, spec, center, inv_ls):
        if self._bits:
            words, popc, ok = self.ops.pack_bits(X)
            if not ok:
                raise ValueError("non-binary rows on a bit-packed path")
            return PointSet(words, words.stride(0), popc, 1, X.shape[0], X.shape[1])
        if self._use_records(spec, X.shape[1]):
            return self.ops.make_records(X, center, inv_ls)
        return self.ops.prepare_points(X, center, inv_ls) if spec.stationary else self.ops.raw_points(X)

    def _gram_T(self, pointset, table):
        """k(table, points)^T as an (m x L) matrix: K1 with one row of m singleton groups and unit weights."""
        m = pointset.n
        at, _ = self.ops.group_accumulate(pointset, table, None, None, m, 0, 0, m, rec=pointset.rec,
                                          unit_weights=True)
        return at

    def _landmarks(self, Z, spec, center, inv_ls):
        """Landmark table over L' = L (+ n_obs) stacked landmarks, and the pieces of the predictive-covariance
        correction (SOBER/_kernel.py:35-53): everything the K1 passes need that does not depend on the basis."""
        lm = {"table": None, "table_z": None, "k_oz": None, "k_zo_w": None, "m_z": None, "table_obs": None}
        if spec is not None:
            lm["table_z"] = lm["table"] = self._table(Z, spec, center, inv_ls)
            if spec.posterior:
                x_obs = self.ops.f64(spec.x_obs)
                w = self.ops.f64(spec.woodbury)
                lm["k_oz"] = self._gram_T(self._points(x_obs, spec, center, inv_ls), lm["table_z"])   # (n_obs x L)
                lm["k_zo_w"] = lm["k_oz"].T @ w                                                       # (L x n_obs)
                if not spec.gspace:
                    lm["table"] = self._table(torch.cat([Z, x_obs], 0), spec, center, inv_ls)
                if spec.weighted or spec.gspace:
                    # predictive mean at the landmarks, m(z) = c + k(z, X_obs) alpha (SOBER/_gp.py:240-253)
                    lm["m_z"] = spec.mean_const + lm["k_oz"].T @ self.ops.f64(spec.alpha)
                    lm["table_obs"] = self._table(x_obs, spec, center, inv_ls)
                if spec.gspace:
                    # mu_g(z) = exp(mu_h + var_h / 2) - 1 (SOBER/BASQ/_scale_mmlt.py:208-220); var_h includes the noise
                    var_z = (spec.outputscale - (lm["k_zo_w"] * lm["k_oz"].T).sum(1) + spec.noise).clamp_min(1e-10)
                    lm["m_z"] = torch.expm1(lm["m_z"] + 0.5 * var_z)
        return lm

    def _basis(self, Z, n_basis, kernel, spec, center, inv_ls, lm):
        """-> U (n x L), Uext (n x L')."""
        o = self.opts
        k_zo_w = lm["k_zo_w"]
        # projector null spaces see the basis only through its row space: skip the final rotation (lowrank_basis)
        rotate = o.rotate_basis if o.rotate_basis is not None else not (o.nullspace == "projector" and self.nullspace is None)
        if self.basis is not None:
            U = self.ops.f64(self.basis)
        elif spec is not None and o.gram == "cuda":
            gram = self._gram_T(self._points(Z, spec, center, inv_ls), lm["table_z"])        # (L x L)
            if k_zo_w is not None:
                gram = gram - k_zo_w @ lm["k_oz"]
            if spec.gspace:
                gram = torch.expm1(gram)
            if lm["m_z"] is not None:
                gram = lm["m_z"].unsqueeze(1) * gram * lm["m_z"].unsqueeze(0)
            if spec.gspace and spec.jitter != 0.0:
                gram.diagonal().add_(spec.jitter)
            gram = 0.5 * (gram + gram.T)
            if o.gate == "cholesky" and o.defer_gate and o.nystrom_qr != "householder":
                # the gate's L x L Cholesky only decides: it runs beside the range finder, which speculates on "passed"
                gram, test = _psd.repair_deferred(gram)
                probe = _nystrom.draw_test_matrix(gram.shape[-1], n_basis, gram.dtype, gram.device)
                U = _nystrom.lowrank_basis(gram, n_basis, qr=o.nystrom_qr, rotate=rotate, probe=probe)
                if not test.passed():
                    gram = _psd.escalate(gram, o.gate)
                    U = _nystrom.lowrank_basis(gram, n_basis, qr=o.nystrom_qr, rotate=rotate, probe=probe)
            else:
                gram = _psd.repair(gram, o.gate, assume_asymmetric=True)
                U = _nystrom.lowrank_basis(gram, n_basis, qr=o.nystrom_qr, rotate=rotate)
        else:
            gram = kernel(Z, Z)
            gram = _psd.repair(gram, o.gate)
            U = _nystrom.lowrank_basis(gram, n_basis, qr=o.nystrom_qr, rotate=rotate)
        Um = U if lm["m_z"] is None else U * lm["m_z"].unsqueeze(0)       # weighted mode: the basis carries m(z_l)
        stacked = k_zo_w is not None and not (spec is not None and spec.gspace)      # linear posterior modes only
        Uext = torch.cat([Um, -(Um @ k_zo_w)], 1) if stacked else Um
        return U, Uext.contiguous()

    def _nystrom(self, Z, n_basis, kernel, spec, center, inv_ls):
        """-> U (n x L), Uext (n x L'), landmark table."""
        lm = self._landmarks(Z, spec, center, inv_ls)
        U, Uext = self._basis(Z, n_basis, kernel, spec, center, inv_ls, lm)
        return U, Uext, lm["table"]

    # -----------------------------------------------------------------------------------------------------
    # one K1 pass over the local alive-list (fused kernel or generic callable)
    # -----------------------------------------------------------------------------------------------------
    def _accumulate(self, st, alive, n_local, pos0, ES, S, unit=False):
        idx, mu = alive.idx, (None if unit else alive.mass)
        if st["spec"] is not None:
            if st.get("post") is not None:
                return self.ops.group_accumulate(st["pts"], st["table"], idx, mu, n_local, pos0, ES, S, post=st["post"])
            return self.ops.group_accumulate(st["pts"], st["table"], idx, mu, n_local, pos0, ES, S, rec=alive.rec,
                                             unit_weights=unit)
        L = st["Z"].shape[0]
        dev = self.ops.device
        at = torch.zeros((S, L), dtype=torch.float64, device=dev)
        totw = torch.zeros(S, dtype=torch.float64, device=dev)
        step = max(int(self.opts.generic_chunk), 1)
        for c0 in range(0, n_local, step):
            c1 = min(n_local, c0 + step)
            rows = st["X"][idx[c0:c1].long()] if idx is not None else st["X"][c0:c1]
            tile = st["kernel"](st["Z"], rows).to(torch.float64).contiguous()
            self.ops.group_accumulate_gram(tile, None if mu is None else mu[c0:c1], pos0 + c0, ES, S, at, totw)
        return at, totw

    # -----------------------------------------------------------------------------------------------------
    def run(self, pts_rec, pts_nys, num_pts, kernel, init_weights=None, calc_obj=None):
        ops, comm, o = self.ops, self.comm, self.opts
        dev = ops.device
        # Candidates in pinned host memory (the end-to-end path): the copy runs in row chunks on a side stream and the
        # first K1 pass consumes the chunks as they land (record layout, plain kernel mode); every other path simply
        # waits for the whole copy.
        incoming = None
        if (torch.is_tensor(pts_rec) and pts_rec.device.type == "cpu" and pts_rec.dtype == torch.float64
                and pts_rec.dim() == 2 and pts_rec.is_contiguous() and pts_rec.is_pinned() and dev.type == "cuda"
                and o.overlap and o.stats is None and calc_obj is None and hasattr(ops, "upload_chunks")
                and pts_rec.shape[0] >= (1 << 18)):
            X, incoming = ops.upload_chunks(pts_rec)
        else:
            X = ops.f64(pts_rec)
        Z = ops.f64(pts_nys)
        if X.dim() != 2 or Z.dim() != 2 or X.shape[1] != Z.shape[1]:
            raise ValueError("pts_rec (N, d) and pts_nys (L, d) must be 2-D with the same d")
        n_rows = X.shape[0]
        counts = comm.all_gather_ints(n_rows, dev)
        row0, n_total = sum(counts[:comm.rank]), sum(counts)
        if n_total >= 2 ** 31:
            raise ValueError("sober_b200 supports fewer than 2^31 candidates")

        if init_weights is None:
            mu = torch.full((n_rows,), 1.0, dtype=torch.float64, device=dev) / n_total
            clearing = None
        else:
            clearing = None
            if init_weights.shape != (n_rows,):
                raise ValueError("init_weights must have shape (len(pts_rec),)")
            mu = ops.f64(init_weights)
            if init_weights.device.type == "cpu" and dev.type == "cuda" and n_rows >= (1 << 18):
                # the caller's HOST vector ends up all zero but for <= b entries: clear it beside the GPU work instead
                # of after it (80 MB of memset at 1e7 candidates), once the upload above has read it
                uploaded = torch.cuda.Event()
                uploaded.record(torch.cuda.current_stream(dev))

                def _clear():
                    uploaded.synchronize()
                    init_weights.zero_()
                clearing = threading.Thread(target=_clear, daemon=True)
                clearing.start()

        clock = _StageClock(o.stats, dev, getattr(o, "nvtx", False))
        spec = introspect(kernel) if o.fuse else None
        if spec is not None and spec.d is not None and spec.d != X.shape[1]:
            raise ValueError("lengthscale dimension does not match the inputs")
        center = inv_ls = None
        d = X.shape[1]
        if spec is not None and spec.stationary:
            # the family constant rides on the lengthscale: the kernels see exp(-d2) / f(sqrt(d2))  (_lib.FAMILY_SCALE)
            center = Z.mean(0).contiguous()
            inv_ls = (ops.f64(spec.inv_ls) * _lib.FAMILY_SCALE[spec.family]).expand(d).contiguous()
        elif spec is not None:
            center = torch.zeros(d, dtype=torch.float64, device=dev)
            inv_ls = torch.ones(d, dtype=torch.float64, device=dev)
        # {0,1}-valued inputs: rows are bit-packed once (64x less data) and the contraction runs on the integer pipe.
        #   Tanimoto: <x, z> = popcount(x & z);
        #   stationary kernel with a single lengthscale (examples/ising.py: RBF on {0,1}^24): the squared distance is
        #   the Hamming distance times a constant, so the kernel value is a (d + 1)-entry lookup table.
        self._bits, self._lut = False, None
        cand_bits = None
        want = None
        if (spec is not None and not spec.gspace and _lib.RECORD_MAX_D < d <= _lib.BITS_MAX_D and o.k1_variant != 1
                and hasattr(ops, "pack_bits")):
            if spec.family == _lib.TANIMOTO:
                want = "tanimoto"
            elif spec.stationary and spec.inv_ls.numel() == 1:
                want = "hamming"
        if want is not None:
            zw, zp, z_ok = ops.pack_bits(Z)
            if z_ok and (spec.x_obs is None or ops.pack_bits(ops.f64(spec.x_obs))[2]):
                xw, xp, x_ok = ops.pack_bits(X)
                flags = comm.all_gather_ints(int(x_ok), dev)
                if all(flags):
                    self._bits = want
                    cand_bits = PointSet(xw, xw.stride(0), xp, 1, X.shape[0], d)
                    if want == "hamming":
                        step = (float(spec.inv_ls.reshape(-1)[0]) * _lib.FAMILY_SCALE[spec.family]) ** 2
                        h = torch.arange(d + 1, dtype=torch.float64, device=dev)
                        self._lut = _family_values(spec.family, h * step).contiguous()
        records = self._use_records(spec, d) and not self._bits
        if incoming is not None and not (records and spec.mode == "kernel" and self.trace is None):
            torch.cuda.current_stream(dev).wait_event(incoming[-1][2])
            incoming = None

        clock.lap("setup")
        lm = self._landmarks(Z, spec, center, inv_ls)
        n = min(num_pts - 1, Z.shape[0]) if self.basis is None else self.basis.shape[0]
        S = 2 * (n + 1)
        st = {"spec": spec, "table": lm["table"], "kernel": kernel, "X": X, "Z": Z,
              "pts": (cand_bits if cand_bits is not None else self._points(X, spec, center, inv_ls))
              if (spec is not None and not records) else None}

        idx, mass, n_local = ops.compact_nonzero(mu)
        # every weight non-zero (the usual case): the alive-list is the identity and the record pass reads X as one
        # contiguous stream instead of gathering rows
        gather = None if n_local == n_rows else idx
        if incoming is not None and (gather is not None or n_local <= S):
            torch.cuda.current_stream(dev).wait_event(incoming[-1][2])
            incoming = None
        if incoming is not None:
            alive = Alive(idx, mass, torch.empty((n_rows, record_stride(d)), dtype=torch.float64, device=dev))
            alive.pending = incoming
        else:
            alive = Alive(idx, mass, ops.make_records(X, center, inv_ls, gather, mass).rec if records else None)
        live = comm.all_gather_ints(n_local, dev)
        pos0, remaining = sum(live[:comm.rank]), sum(live)
        obj = None if calc_obj is None else (-1 * calc_obj(pts_rec.to(dev))).to(torch.float64)
        clock.lap("compact+records")

        # weighted mode (SOBER/_kernel.py:33-47): predictive mean m(x) of every local candidate, once;
        # gspace mode (SOBER/BASQ/_scale_mmlt.py:208-220, 256-275): mu_g(x) = exp(mu_h + var_h / 2) - 1, and the rows
        # k(x, X_obs) are kept: the POST mode of K1 contracts them with k(z, X_obs) W inside the kernel
        m_x = None
        if spec is not None and (spec.weighted or spec.gspace):
            alpha = ops.f64(spec.alpha)
            m_x = torch.empty(n_rows, dtype=torch.float64, device=dev)
            n_obs = alpha.numel()
            kx_all = torch.empty((n_rows, n_obs), dtype=torch.float64, device=dev) if spec.gspace else None
            wood = ops.f64(spec.woodbury) if spec.gspace else None
            for c0 in range(0, n_rows, 1 << 17):
                c1 = min(n_rows, c0 + (1 << 17))
                chunk = X[c0:c1]
                pts_c = self.ops.prepare_points(chunk, center, inv_ls) if (spec.gspace and spec.stationary) \
                    else self._points(chunk, spec, center, inv_ls)
                rows = self._gram_T(pts_c, lm["table_obs"])                                  # (m x n_obs)
                mean = spec.mean_const + rows @ alpha
                if spec.gspace:
                    kx_all[c0:c1] = rows
                    var = (spec.outputscale - ((rows @ wood) * rows).sum(1) + spec.noise).clamp_min(1e-10)
                    mean = torch.expm1(mean + 0.5 * var)
                m_x[c0:c1] = mean
            if spec.gspace:
                st["post"] = (kx_all, lm["k_zo_w"].contiguous())
        self._m_x = m_x

        fast_tail = (comm.world == 1 and obj is None and m_x is None and not o.fused_projection
                     and hasattr(ops, "apply_tail") and dev.type == "cuda")

        def k1_pass(alive, n_local, pos0, remaining):
            """Grouped kernel-column sums of one iteration: At, totw and the second count of the remainder."""
            ES = (remaining // S) * S
            t0 = min(max(ES - pos0, 0), n_local)           # first local offset belonging to the remainder
            src = alive
            if m_x is not None:
                # the kernel columns are weighted with mu_i m(x_i); the group MASSES stay sums of mu_i
                wts = alive.mass * m_x[alive.idx.long()]
                if alive.rec is not None:
                    alive.rec[:, d + 1] = wts
                src = Alive(alive.idx, wts, alive.rec)
            if alive.pending is not None:
                # pipelined upload: build the records of a chunk and add its group sums as soon as its rows have landed
                # (a chunk is handled like a rank's shard: positions [pos0 + a, pos0 + b))
                at = totw = None
                stream = torch.cuda.current_stream(dev)
                for a, b, landed in alive.pending:
                    stream.wait_event(landed)
                    ops.make_records(X[a:b], center, inv_ls, None, alive.mass[a:b], out=alive.rec[a:b])
                    at_c, tw_c = self._accumulate(st, alive.part(a, b), b - a, pos0 + a, ES, S)
                    at, totw = (at_c, tw_c) if at is None else (at.add_(at_c), totw.add_(tw_c))
                alive.pending = None
            else:
                at, totw = self._accumulate(st, src, n_local, pos0, ES, S)
            if m_x is not None:
                lead = pos0 % S
                padded = torch.zeros(-(-(lead + t0) // S) * S, dtype=torch.float64, device=dev)
                padded[lead:lead + t0] = alive.mass[:t0]
                totw = padded.reshape(-1, S).sum(0)
            Lp = at.shape[1]
            if fast_tail:
                # single process, no objective: the second count is applied by ONE kernel (ops.apply_tail) instead of
                # being packed into ``extra`` for the all-reduce and added with five small torch launches
                tail = None
                if t0 < n_local:
                    tail_at, tail_tw = self._accumulate(st, src.tail(t0), n_local - t0, 0, n_local - t0, 1)
                    tail = (tail_at[0], tail_tw)
                return at, totw, tail, t0
            extra = torch.zeros(Lp + 3, dtype=torch.float64, device=dev)
            if t0 < n_local:
                # second count of the remainder into the last group (SOBER/_rchq.py:153-164)
                tail_at, tail_tw = self._accumulate(st, src.tail(t0), n_local - t0, 0, n_local - t0, 1)
                extra[:Lp] = tail_at[0]
                extra[Lp] = tail_tw[0] if m_x is None else alive.mass[t0:].sum()
            return at, totw, extra, t0

        # The first (and by far the longest) K1 pass does not depend on the Nystrom basis: it runs on a stream confined
        # to all SMs but a few (a green context) beside the range finder -- a host-driven sequence of ~300 small
        # kernels (Cholesky-QR passes, eigh, refinement sweeps) that leaves most of the GPU idle.
        first = {}
        fork = contextlib.nullcontext()
        if (remaining > S and o.overlap and o.stats is None and self.trace is None and dev.type == "cuda"
                and self.basis is None and hasattr(ops, "partition_stream")):
            part = ops.partition_stream()
            if part is not None:
                def launch_first():
                    first["k1"] = k1_pass(alive, n_local, pos0, remaining)
                    at_, totw_, tail_ = first["k1"][:3]
                    made = [at_, totw_]
                    if tail_ is not None:
                        made += list(tail_) if isinstance(tail_, tuple) else [tail_]
                    return made
                fork = _nystrom.SideStream(launch_first, dev, part)
        with fork:
            U, Uext = self._basis(Z, n, kernel, spec, center, inv_ls, lm)
        if U.shape[0] != n:                     # fewer landmarks than basis functions: the basis decides
            n = U.shape[0]
            S = 2 * (n + 1)
            first.clear()
        clock.lap("nystrom")
        UextT = Uext.T.contiguous()

        while True:
            if remaining <= S:
                sel_idx, sel_w = self._finish(st, alive, n_local, pos0, remaining, n, UextT, row0, obj, n_rows)
                clock.lap("finish")
                break
            E = remaining // S
            ES = E * S
            idx, mass = alive.idx, alive.mass
            at, totw, extra, t0 = first.pop("k1") if "k1" in first else k1_pass(alive, n_local, pos0, remaining)
            clock.lap("k1")
            Lp = at.shape[1]
            objs = None
            if obj is not None:
                objs = torch.zeros((S, 1), dtype=torch.float64, device=dev)
                if n_local > 0:
                    row = obj[idx.long()].reshape(1, -1).contiguous()
                    ops.group_accumulate_gram(row, mass, pos0, ES, S, objs, None)
                    if t0 < n_local:
                        extra[Lp + 1] = torch.dot(row[0, t0:], mass[t0:])
            if comm.world > 1:
                # ONE all-reduce per iteration: group sums, group masses, the remainder's second count and -- with an
                # objective (SOBER/_rchq.py:138-150) -- the per-group objective sums ride in the same packed buffer
                parts = [at.reshape(-1), totw, extra] + ([objs.reshape(-1)] if objs is not None else [])
                packed = torch.cat(parts)
                t_ar = ops._begin("all_reduce") if hasattr(ops, "_begin") else None
                comm.all_reduce(packed)
                if t_ar is not None:
                    ops._end("all_reduce", t_ar, packed.numel() * 8)
                at = packed[:S * Lp].reshape(S, Lp)
                totw = packed[S * Lp:S * Lp + S]
                extra = packed[S * Lp + S:S * Lp + S + Lp + 3]
                if objs is not None:
                    objs = packed[S * Lp + S + Lp + 3:].reshape(S, 1)
            # The fused DMMA projection kernel (csrc/project.cu) is correct but was measured 7x slower than the cuBLAS
            # DGEMM + 4 elementwise ops it replaces (0.14 ms vs 0.02 ms per call at C2); it stays opt-in.
            fused = (o.fused_projection and obj is None and self.trace is None and hasattr(ops, "project_design"))
            if fused:
                # projection + second count of the remainder + barycentres + ones column: one DMMA kernel
                design, totw = ops.project_design(at, Uext, totw, tail=extra[:Lp], tail_tw=extra[Lp:Lp + 1])
                bary = None
            else:
                if fast_tail:
                    totw = ops.apply_tail(at, totw, *(extra if extra is not None else (None, None)))
                else:
                    at[S - 1] += extra[:Lp]
                    totw = totw.clone()
                    totw[S - 1] += extra[Lp]
                bary = at @ UextT                                   # (S x n)  == (U @ X_for_nys).T
                if objs is not None:
                    objs[S - 1, 0] += extra[Lp + 1]
                    bary = torch.cat([bary, objs], 1)
                if self.trace is not None:
                    self.trace("group", {"At": at.clone(), "totw": totw.clone(), "Xt_unnormalised": bary.clone(),
                                         "R": remaining, "E": E})
                design = None
            graph_step = (design is None and self.nullspace is None and o.nullspace == "projector" and objs is None
                          and self.trace is None)
            if design is None and not graph_step:
                bary = bary / totw.unsqueeze(1)
            clock.lap("tail+project")
            rank = keep = early = None
            n_design = (design.shape[1] if design is not None else bary.shape[1] + 1)
            if graph_step:
                # the whole step (barycentres, null space, elimination, survivor counts and ranks) as one CUDA-graph replay
                wfull, kept, summary, rank = _car.reduce_step(ops, bary, totw, use_graph=o.graphs and o.stats is None,
                                                              divide=True, comm=comm)
                # the weight update + compaction is enqueued BEFORE the host reads the survivor counts: the kernel takes
                # them from the device summary (closed form of KeepMap.before), so the GPU goes on while the host syncs
                # and prepares the next K1 launch.  Discarded in the (rare) retry case.
                if hasattr(ops, "update_compact_dev") and n_local > 0:
                    # ... and the counts travel to the host on a side stream that waits for the Caratheodory step only,
                    # not for the update kernel behind it
                    fetch = ops.fetch_small(summary) if hasattr(ops, "fetch_small") else None
                    early = ops.update_compact_dev(idx, mass, n_local, pos0, ES, S, wfull, totw, rank, summary,
                                                   rec=alive.rec, d=d)
                    summary = fetch() if fetch is not None else summary.tolist()
                else:
                    summary = summary.tolist()                                           # the one host sync of the iteration
                keep = KeepMap.from_summary(summary, S, ES)
                retry = _car.needs_retry("projector", keep.K, n_design, bool(summary[S]))
                if retry:
                    early = None
                    bary = bary / totw.unsqueeze(1)
            else:
                wfull = _car.caratheodory(ops, bary, totw, o.nullspace, self.nullspace, design=design)
                kept = wfull > 0
                flags = torch.cat([kept, torch.isfinite(wfull).all().reshape(1)]).tolist()   # the one host sync
                retry = self.nullspace is None and _car.needs_retry(o.nullspace, sum(flags[:-1]), n_design, flags[-1])
            if retry:
                if o.stats is not None:
                    o.stats["car_retries"] = o.stats.get("car_retries", 0) + 1
                wfull = _car.caratheodory(ops, bary, totw, "qr", design=design)
                kept = wfull > 0
                flags = kept.tolist() + [True]
                rank = keep = None
            clock.lap("car")
            if obj is not None:
                wfull = self._objective_step(bary[:, :n], bary[:, n], wfull)
                kept = wfull > 0
                flags = kept.tolist() + [True]
                rank = keep = None
            if rank is None:
                rank = (torch.cumsum(kept.to(torch.int32), 0) - kept.to(torch.int32)).to(torch.int32)
            if keep is None:
                keep = KeepMap(flags[:-1], S, ES)
            new_pos0 = keep.before(pos0)
            new_local = keep.before(pos0 + n_local) - new_pos0
            if early is not None:
                alive = Alive(early[0][:new_local], early[1][:new_local],
                              None if early[2] is None else early[2][:new_local])
            else:
                alive = Alive(*ops.update_compact(idx, mass, n_local, pos0, ES, S, wfull, totw, rank, keep.K,
                                                  keep.tail_keep, new_pos0, new_local, rec=alive.rec, d=d))
            pos0, n_local, remaining = new_pos0, new_local, keep.before(remaining)
            clock.lap("keepmap+update")

        # in-place sparse result in the caller's weight vector (SOBER/_rchq.py:109-110, 203-218)
        if init_weights is not None:
            mine = (sel_idx >= row0) & (sel_idx < row0 + n_rows)
            loc, w_loc = (sel_idx[mine] - row0).contiguous(), sel_w[mine].contiguous()
            if mu.data_ptr() == init_weights.data_ptr():
                ops.scatter_result(mu, loc, w_loc)
            else:
                if clearing is not None:
                    clearing.join()
                else:
                    init_weights.zero_()
                init_weights[loc.to(init_weights.device)] = w_loc.to(init_weights.device, init_weights.dtype)
            sel_w = sel_w.to(init_weights.dtype)
        return sel_idx, sel_w

    # -----------------------------------------------------------------------------------------------------
    def _finish(self, st, alive, n_local, pos0, remaining, n, UextT, row0, obj, n_rows=0):
        """The two terminal branches, SOBER/_rchq.py:72-75 (R <= n+1) and :77-114 (n+1 < R <= S)."""
        ops, comm, o = self.ops, self.comm, self.opts
        dev = ops.device
        idx, mass = alive.idx, alive.mass
        all_mass = torch.zeros(remaining, dtype=torch.float64, device=dev)
        all_idx = torch.zeros(remaining, dtype=torch.int64, device=dev)
        all_mass[pos0:pos0 + n_local] = mass
        all_idx[pos0:pos0 + n_local] = idx.long() + row0
        if remaining <= n + 1:
            if comm.world > 1:
                comm.all_reduce(all_mass)
                comm.all_reduce(all_idx)
            live = all_mass > 0
            return all_idx[live], all_mass[live]
        if remaining > 0 and n_local > 0:
            feats_t, _ = self._accumulate(st, alive, n_local, pos0, 0, remaining, unit=True)   # (R x L'), unit weights
        else:
            Lp = UextT.shape[0]
            feats_t = torch.zeros((remaining, Lp), dtype=torch.float64, device=dev)
        if comm.world > 1:
            comm.all_reduce(feats_t)
            comm.all_reduce(all_mass)
            comm.all_reduce(all_idx)
        if getattr(self, "_m_x", None) is not None:
            # weighted mode: feature rows carry m(x_i) (the basis already carries m(z_l))
            all_m = torch.zeros(remaining, dtype=torch.float64, device=dev)
            all_m[pos0:pos0 + n_local] = self._m_x[idx.long()]
            if comm.world > 1:
                comm.all_reduce(all_m)
            feats_t = feats_t * all_m.unsqueeze(1)
        feats = feats_t @ UextT                                                            # (R x n)
        head_obj = None
        if obj is not None:
            # objective values of the remaining points, and -- for the reference's position-indexed lookup below -- of
            # the first ``remaining`` ROWS of the global candidate set; both assembled from the row shards
            alive_obj = torch.zeros(remaining, dtype=torch.float64, device=dev)
            alive_obj[pos0:pos0 + n_local] = obj[idx.long()]
            head_obj = torch.zeros(remaining, dtype=torch.float64, device=dev)
            lo, hi = min(row0, remaining), min(row0 + (n_rows or obj.numel()), remaining)
            if hi > lo:
                head_obj[lo:hi] = obj[lo - row0:hi - row0]
            if comm.world > 1:
                comm.all_reduce(alive_obj)
                comm.all_reduce(head_obj)
            feats = torch.cat([feats, alive_obj.unsqueeze(1)], 1)
        if self.nullspace is None and o.nullspace == "projector":
            wfull, _, summary, _ = _car.reduce_step(ops, feats, all_mass,
                                                    use_graph=o.graphs and o.stats is None and self.trace is None,
                                                    comm=comm)
            summary = summary.tolist()
            if _car.needs_retry("projector", summary[-2], feats.shape[1] + 1, bool(summary[-1])):
                wfull = _car.caratheodory(ops, feats, all_mass, "qr")
        else:
            wfull = _car.caratheodory(ops, feats, all_mass, o.nullspace, self.nullspace)
        if obj is not None:
            # NB the reference indexes ``obj`` with POSITIONS here (SOBER/_rchq.py:89), kept as is
            live = torch.nonzero(wfull > 0).reshape(-1)
            wfull = self._objective_step(feats[:, :n], None, wfull, obj_vals=head_obj[live])
        live = wfull > 0
        return all_idx[live], wfull[live]

    def _objective_step(self, feats, obj_col, wfull, obj_vals=None):
        """One extra null-space move on the n+2 survivors, signed to increase the objective
        (SOBER/_rchq.py:87-106 and :177-196).  Tiny: runs as torch ops on the device."""
        live = torch.nonzero(wfull > 0).reshape(-1)
        w = wfull[live]
        pts = torch.cat([feats[live].T, torch.ones((1, len(live)), dtype=feats.dtype, device=feats.device)], 0)
        _, _, vh = torch.linalg.svd(pts)
        direction = vh[-1]
        vals = obj_col[live] if obj_vals is None else obj_vals
        if torch.dot(vals, direction) < 0:
            direction = -direction
        pos = direction > 0
        ratio = torch.zeros_like(w)
        ratio[pos] = w[pos] / direction[pos]
        cand = torch.arange(len(w), device=w.device)[pos]
        pivot = cand[torch.argmin(ratio[pos])]
        w = w - ratio[pivot] * direction
        w[pivot] = 0.0
        out = torch.zeros_like(wfull)
        out[live] = torch.where(w > 0, w, torch.zeros_like(w))
        return out


# ---------------------------------------------------------------------------------------------------------
# public entry point
# ---------------------------------------------------------------------------------------------------------
_default_ops = {}      # one CudaOps (workspaces, graph caches, partition stream) per CUDA device index
_default_comm = None


def _ops():
    from ._ops import CudaOps
    key = torch.cuda.current_device() if torch.cuda.is_available() else -1
    ops = _default_ops.get(key)
    if ops is None:
        ops = _default_ops[key] = CudaOps()
    ops.variant = options.k1_variant
    return ops


def set_communicator(comm):
    """``None`` -> single process; a ``Sharded`` instance -> row-sharded candidates."""
    global _default_comm
    _default_comm = comm


def recombination(
    pts_rec,            # candidates (N, d); with a communicator set: this rank's contiguous row shard
    pts_nys,            # Nystrom landmarks (L, d), replicated
    num_pts,            # batch size b: at most b points are returned
    kernel,             # SOBER Kernel object or any callable kernel(x, y) -> Gram
    device,             # ignored, as in the reference (SOBER/_rchq.py:30): the CUDA device is used
    dtype,              # ignored, as in the reference: float64 arithmetic
    init_weights=None,  # (N,) importance weights; MUTATED IN PLACE into the sparse solution
    calc_obj=None,      # optional objective callable (SOBER/_rchq.py:67-69)
):
    """Same contract as ``SOBER._rchq.recombination`` (SOBER/_rchq.py:5-31): returns ``(idx, w)`` with ``idx`` the
    ascending int64 indices of at most ``num_pts`` selected candidates and ``w > 0`` their weights
    (``w.sum() == init_weights.sum()``), both on the CUDA device."""
    return Recombiner(_ops(), _default_comm).run(pts_rec, pts_nys, num_pts, kernel, init_weights, calc_obj)
