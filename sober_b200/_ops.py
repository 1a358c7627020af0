"""Device operations of the recombination loop: torch tensors in, C-ABI calls out.

``CudaOps`` is the only implementation shipped.  It requires a CUDA device and the built extension and raises
otherwise (no CPU fallback).  Tensors are allocated by torch (caching allocator), kernels are enqueued on
torch's current stream; nothing here synchronises except where a count has to reach the host.
"""
import contextlib
import ctypes as C
import os

import torch

from . import _lib
from ._lib import GroupArgs, check

_FAMILY_NAMES = {"rbf": _lib.RBF, "matern12": _lib.MATERN12, "matern32": _lib.MATERN32,
                 "matern52": _lib.MATERN52, "tanimoto": _lib.TANIMOTO}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class PointSet:
    """Candidate (or landmark-as-candidate) rows in a layout K1 reads.

    indexed layout: ``rows`` (n x ld) + ``xn`` view, addressed through an alive-list of row ids;
    record layout : ``rec`` (m x ldr) rows ``[coords | norm | weight | pad]`` already in alive-list order."""

    def __init__(self, rows, ld, xn, xn_stride, n, d, rec=None, ldr=0):
        self.rows, self.ld, self.xn, self.xn_stride, self.n, self.d = rows, ld, xn, xn_stride, n, d
        self.rec, self.ldr = rec, ldr


def record_stride(d):
    return (d + 3) // 2 * 2


class LandmarkTable:
    """``Zt`` (L x d) and ``zn`` (L) as K1 wants them, plus the family / output scale.  ``d`` overrides the column
    count when the rows are bit-packed words (family TANIMOTO_BITS: d = number of bits)."""

    def __init__(self, zt, zn, family, outputscale, d=None, lut=None):
        self.zt, self.zn, self.family, self.outputscale = zt, zn, family, float(outputscale)
        self.lut = lut          # family HAMMING_LUT: kernel value per Hamming distance (d + 1 doubles)
        self.L = zt.shape[0]
        self.d = zt.shape[1] if d is None else int(d)


_NO_GUARD = contextlib.nullcontext()

# raw equivalents of torch.cuda.current_stream(i).cuda_stream / torch.cuda.current_device() (private but stable since
# torch 1.x; the public calls are used if a build lacks them)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda i: torch.cuda.current_stream(i).cuda_stream)
_current_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


class CudaOps:
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.SoberB200Error(
                "sober_b200 needs a CUDA device (B200, sm_100a); there is no CPU implementation of the hot path")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.variant = 0
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ws = {}
        self.launches = 0      # kernels launched through the C ABI (bench.py reports it)
        self.timing = None     # None, or {name: [(start_event, end_event, work), ...]} filled per call

    # -- helpers -------------------------------------------------------------------------------------
    # The wrappers below run a dozen times per loop iteration between a host sync and the next big kernel, i.e. with the
    # GPU idle: torch.cuda.current_stream() (8 us: builds a Stream object) and the torch.cuda.device() guard (7 us)
    # are replaced by their raw equivalents.
    def _stream_ptr(self):
        return _raw_stream(self.index)

    def _stream(self):
        return C.c_void_p(_raw_stream(self.index))

    def _guard(self):
        """Device guard that costs nothing when the current device already is ours (one process per GPU)."""
        if _current_device() == self.index:
            return _NO_GUARD
        return torch.cuda.device(self.device)

    def _workspace(self, nbytes):
        # one workspace per stream: the first K1 pass runs on the SM-partitioned stream beside work on the main one
        key = self._stream_ptr()
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = self._ws[key] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=self.device)
        return ws

    def f64(self, t):
        return t.to(device=self.device, dtype=torch.float64, non_blocking=True).contiguous()

    def _begin(self, name):
        if self.timing is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def _end(self, name, start, work):
        if start is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self.timing.setdefault(name, []).append((start, ev, work))

    def timing_largest(self, name):
        """(ms, work) of the launch with the most work recorded under ``name`` -- call after a synchronize."""
        recs = (self.timing or {}).get(name, [])
        if not recs:
            return None
        a, b, w = max(recs, key=lambda r: r[2])
        return a.elapsed_time(b), w

    def timing_summary(self):
        """{name: (calls, total_ms, total_work)} -- call after a synchronize."""
        out = {}
        for name, recs in (self.timing or {}).items():
            out[name] = (len(recs), sum(a.elapsed_time(b) for a, b, _ in recs), sum(w for _, _, w in recs))
        return out

    # -- streaming passes ------------------------------------------------------------------------------
    def prepare_points(self, X, center, inv_ls):
        """(X - c) * inv_ls with the squared norm appended: the candidate layout for stationary families."""
        n, d = X.shape
        ldp = (d + 1 + 1) // 2 * 2  # even row stride keeps rows 16-byte aligned
        P = torch.empty((n, ldp), dtype=torch.float64, device=self.device)
        with self._guard():
            t0 = self._begin("prepare_points")
            check(self.lib.sober_prepare_points(_ptr(X), X.stride(0), n, d, _ptr(center), _ptr(inv_ls), _ptr(P), ldp,
                                                self._stream()), "prepare_points")
            self._end("prepare_points", t0, 8 * n * (d + ldp))
        self.launches += 1
        return PointSet(P, ldp, P[:, d], ldp, n, d)

    def upload_chunks(self, host, min_rows=1 << 16, max_chunks=8):
        """Host (pinned) -> device copy of a row-major float64 matrix in row chunks on a side stream.  Returns the device
        tensor and [(row_begin, row_end, event)]: a consumer makes its stream wait for a chunk's event before reading the
        rows, so the first K1 pass starts on chunk 0 while the rest is still crossing PCIe."""
        n = host.shape[0]
        out = torch.empty(host.shape, dtype=torch.float64, device=self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(self.device))
        out.record_stream(cs)
        step = max(min_rows, -(-n // max_chunks))
        chunks = []
        with torch.cuda.stream(cs):
            for a in range(0, n, step):
                b = min(n, a + step)
                out[a:b].copy_(host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                chunks.append((a, b, ev))
        return out, chunks

    def make_records(self, X, center, inv_ls, idx=None, mu=None, out=None):
        """Gather + (x - c) * inv_ls + norm + weight into the record layout (one 16-byte aligned row per point)."""
        d = X.shape[1]
        m = X.shape[0] if idx is None else idx.numel()
        ldr = record_stride(d)
        rec = torch.empty((m, ldr), dtype=torch.float64, device=self.device) if out is None else out
        with self._guard():
            t0 = self._begin("make_records")
            check(self.lib.sober_make_records(_ptr(X), X.stride(0), d, _ptr(center), _ptr(inv_ls), _ptr(idx), _ptr(mu),
                                              m, _ptr(rec), ldr, self._stream()), "make_records")
            self._end("make_records", t0, 8 * m * (d + 2 + ldr))
        self.launches += 1
        return PointSet(None, 0, None, 0, m, d, rec=rec, ldr=ldr)

    def pack_bits(self, X):
        """{0,1}-valued rows -> (words (n x W int64), popcounts (n,), all_binary: bool).  One streaming pass and one
        host sync (the flag).  W = ceil(d / 64) rounded up to a power of two (what the popcount kernel takes)."""
        n, d = X.shape
        W = 1
        while W * 64 < d:
            W *= 2
        words = torch.empty((n, W), dtype=torch.int64, device=self.device)
        popc = torch.empty(n, dtype=torch.float64, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        with self._guard():
            t0 = self._begin("pack_bits")
            check(self.lib.sober_pack_bits(_ptr(X), X.stride(0), n, d, _ptr(words), W, _ptr(popc), _ptr(flag),
                                           self._stream()), "pack_bits")
            self._end("pack_bits", t0, 8 * n * (d + W + 1))
        self.launches += 1
        return words, popc, int(flag.item()) == 0

    def raw_points(self, X):
        """Tanimoto layout: the rows as they are plus |x|^2."""
        n, d = X.shape
        xn = torch.empty(n, dtype=torch.float64, device=self.device)
        with self._guard():
            check(self.lib.sober_row_sqnorm(_ptr(X), X.stride(0), n, d, _ptr(xn), self._stream()), "row_sqnorm")
        self.launches += 1
        return PointSet(X, X.stride(0), xn, 1, n, d)

    def compact_nonzero(self, mu):
        """``arange(N)[mu != 0]`` and the matching weights; returns (idx int32, mu, count) -- one host sync."""
        n = mu.numel()
        idx = torch.empty(n, dtype=torch.int32, device=self.device)
        out = torch.empty(n, dtype=torch.float64, device=self.device)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.device)
        nbytes = self.lib.sober_compact_workspace(n)
        ws = self._workspace(nbytes)
        with self._guard():
            t0 = self._begin("compact_nonzero")
            check(self.lib.sober_compact_nonzero(_ptr(mu), n, _ptr(idx), _ptr(out), _ptr(cnt), _ptr(ws), ws.numel(),
                                                 self._stream()), "compact_nonzero")
            self._end("compact_nonzero", t0, 8 * n * 2 + 12 * n)
        self.launches += 3
        r = int(cnt.item())
        return idx[:r], out[:r], r

    # -- K1 ------------------------------------------------------------------------------------------------
    def group_accumulate(self, pts, lm, idx, mu, n_local, pos0, ES, S, n_global=None, rec=None, unit_weights=False,
                         post=None):
        """At (S x L) and totw (S) for the positions [pos0, pos0 + n_local) this device owns.  ``rec``: record rows of
        exactly those positions (record layout); otherwise ``pts`` + ``idx`` + ``mu`` (indexed layout)."""
        if int(n_local) == 0:
            # a shard with no alive rows (all-zero weights on this rank) contributes nothing; its 0-row tensors have
            # NULL data pointers, which the C side would read as "indexed layout without X" and reject
            return (torch.zeros((S, lm.L), dtype=torch.float64, device=self.device),
                    torch.zeros(S, dtype=torch.float64, device=self.device))
        At = torch.empty((S, lm.L), dtype=torch.float64, device=self.device)
        totw = torch.empty(S, dtype=torch.float64, device=self.device)
        a = GroupArgs()
        if rec is not None:
            a.rec, a.ldr = rec.data_ptr(), rec.stride(0)
            a.unit_weights = int(bool(unit_weights))
        else:
            a.X, a.ldx = pts.rows.data_ptr(), pts.ld
            a.xn, a.xn_stride = pts.xn.data_ptr(), pts.xn_stride
            a.idx = None if idx is None else idx.data_ptr()
            a.mu = None if mu is None else mu.data_ptr()
        a.n_local, a.pos0, a.ES = int(n_local), int(pos0), int(ES)
        a.n_global = int(pos0 + n_local if n_global is None else n_global)
        a.S, a.L, a.d, a.family = int(S), int(lm.L), int(lm.d), int(lm.family)
        a.outputscale = lm.outputscale
        a.Zt, a.zn = lm.zt.data_ptr(), lm.zn.data_ptr()
        a.lut = None if lm.lut is None else lm.lut.data_ptr()
        if post is not None:
            # non-linear posterior mode: kx (N x n_obs) rows k(x, X_obs) by row id, aw (L x n_obs) rows k(z_l, X_obs) W
            kx, aw = post
            assert rec is None and kx.stride(1) == 1 and aw.is_contiguous() and aw.shape == (lm.L, kx.shape[1])
            a.kx, a.ldkx, a.aw, a.n_obs, a.transform = kx.data_ptr(), kx.stride(0), aw.data_ptr(), kx.shape[1], 1
        a.At, a.totw = At.data_ptr(), totw.data_ptr()
        a.variant = self.variant
        nbytes = self.lib.sober_group_accumulate_workspace(C.byref(a))
        if nbytes < 0:
            raise _lib.SoberB200Error("sober_b200: group_accumulate: bad arguments")
        ws = self._workspace(nbytes)
        with self._guard():
            t0 = self._begin("group_accumulate")
            check(self.lib.sober_group_accumulate(C.byref(a), _ptr(ws), ws.numel(), self._stream()),
                  "group_accumulate")
            self._end("group_accumulate", t0, int(n_local) * int(lm.L))      # work = kernel evaluations (pairs)
        self.launches += 2 if nbytes > 0 else 1
        return At, totw

    def group_accumulate_gram(self, G, mu, pos_begin, ES, S, At, totw):
        L, m = G.shape
        with self._guard():
            check(self.lib.sober_group_accumulate_gram(_ptr(G), G.stride(0), L, m, _ptr(mu), int(pos_begin), int(ES),
                                                       int(S), _ptr(At), _ptr(totw), self._stream()),
                  "group_accumulate_gram")
        self.launches += 1

    # -- CAR -----------------------------------------------------------------------------------------------
    def car_eliminate(self, basis_rows, mass, want_pivots=False, exact=True):
        """Runs the elimination in place on ``mass`` (S,); ``basis_rows`` (k x S, contiguous) is destroyed."""
        k, S = basis_rows.shape
        assert basis_rows.is_contiguous() and mass.is_contiguous()
        piv = torch.empty(max(k, 1), dtype=torch.int32, device=self.device) if want_pivots else None
        steps = torch.zeros(1, dtype=torch.int32, device=self.device) if want_pivots else None
        nbytes = self.lib.sober_car_workspace(k)
        flags = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=self.device)
        with self._guard():
            t0 = self._begin("car_eliminate")
            check(self.lib.sober_car_eliminate(_ptr(basis_rows), k, S, _ptr(mass), int(bool(exact)), _ptr(piv), _ptr(steps),
                                               _ptr(flags),
                                               flags.numel() * 4, self._stream()), "car_eliminate")
            self._end("car_eliminate", t0, k)                                   # work = elimination steps
        self.launches += 1
        return (piv, steps) if want_pivots else None

    def project_design(self, at, uext, totw, tail=None, tail_tw=None):
        """[1 | ((At + tail on the last row) @ Uext^T) / totw'] and totw' (totw plus the tail mass on the last group):
        projection, barycentres and the design matrix of the CAR step in one DMMA kernel."""
        S, Lp = at.shape
        n = uext.shape[0]
        design = torch.empty((S, n + 1), dtype=torch.float64, device=self.device)
        totw_out = torch.empty(S, dtype=torch.float64, device=self.device)
        with self._guard():
            t0 = self._begin("project_design")
            check(self.lib.sober_project_design(_ptr(at), at.stride(0), _ptr(tail), _ptr(totw), _ptr(tail_tw), _ptr(uext),
                                                uext.stride(0), S, Lp, n, _ptr(design), design.stride(0), _ptr(totw_out),
                                                self._stream()), "project_design")
            self._end("project_design", t0, 2 * S * Lp * n)
        self.launches += 1
        return design, totw_out

    def car_cluster_fits(self, S, n_prime, have_basis):
        """Cluster size the fused CAR kernel would use for this shape, 0 if it does not fit in distributed smem."""
        with self._guard():
            return int(self.lib.sober_car_cluster_fits(int(S), int(n_prime), int(bool(have_basis))))

    def car_cluster(self, mass, design=None, basis_rows=None, exact=False):
        """Fused CAR on one thread-block cluster: QR null space + elimination from ``design`` (S x n'), or the
        elimination alone on ``basis_rows`` (k x S).  ``mass`` (S,) is reduced in place."""
        src = design if basis_rows is None else basis_rows
        assert src.is_contiguous() and mass.is_contiguous()
        if basis_rows is None:
            S, n_prime = design.shape
        else:
            S = basis_rows.shape[1]
            n_prime = S - basis_rows.shape[0]
        with self._guard():
            t0 = self._begin("car_cluster")
            check(self.lib.sober_car_cluster(_ptr(design) if basis_rows is None else None, _ptr(basis_rows), int(S),
                                             int(n_prime), _ptr(mass), int(bool(exact)), None, self._stream()),
                  "car_cluster")
            self._end("car_cluster", t0, (S - n_prime) + (0 if basis_rows is not None else 2 * n_prime))
        self.launches += 1

    def car_cols_fits(self, S, k):
        with self._guard():
            return int(self.lib.sober_car_cluster_cols_fits(int(S), int(k)))

    def car_cols(self, basis_rows, mass, exact=False):
        """Elimination on ``basis_rows`` (k x S) with the column-distributed cluster kernel; ``mass`` reduced in place."""
        k, S = basis_rows.shape
        assert basis_rows.is_contiguous() and mass.is_contiguous()
        with self._guard():
            t0 = self._begin("car_cols")
            check(self.lib.sober_car_cluster_cols(_ptr(basis_rows), k, S, _ptr(mass), int(bool(exact)), None,
                                                  self._stream()), "car_cluster_cols")
            self._end("car_cols", t0, k)
        self.launches += 1

    def car_panel_fits(self, S, k):
        """0 = unsupported shape, 1 = the basis is one panel (single cluster kernel), 2 = blocked (panels + GEMM updates)."""
        return int(self.lib.sober_car_panel_fits(int(S), int(k)))

    def car_panel(self, basis_rows, mass, nb_hint=0, want_info=False, prof=None):
        """Elimination on ``basis_rows`` (k x S, destroyed) with the panelled kernels of csrc/car_panel.cu (fused
        arithmetic); ``mass`` reduced in place.  ``want_info``: returns the device int32 pair [stopped, steps]."""
        k, S = basis_rows.shape
        assert basis_rows.is_contiguous() and mass.is_contiguous()
        nbytes = self.lib.sober_car_panel_workspace(S, k)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
        info = torch.empty(2, dtype=torch.int32, device=self.device) if want_info else None
        with self._guard():
            t0 = self._begin("car_panel")
            if prof is None:
                check(self.lib.sober_car_panel(_ptr(basis_rows), k, S, _ptr(mass), int(nb_hint), _ptr(info), _ptr(ws),
                                               ws.numel(), self._stream()), "car_panel")
            else:
                check(self.lib.sober_car_panel_profiled(_ptr(basis_rows), k, S, _ptr(mass), int(nb_hint), _ptr(info),
                                                        _ptr(ws), ws.numel(), _ptr(prof), self._stream()), "car_panel")
            self._end("car_panel", t0, k)
        fits = self.car_panel_fits(S, k)
        nb = 64 if not nb_hint else min(int(nb_hint), 64)
        self.launches += 1 if (fits == 1 and not nb_hint) else 3 * ((k + nb - 1) // nb) - 2
        return info

    def car_prepare(self, feats, div=None):
        """Column-normalised design matrix [1 | feats / div] (S x (n + 1)) in one kernel."""
        S, n = feats.shape
        assert feats.stride(1) == 1
        out = torch.empty((S, n + 1), dtype=torch.float64, device=self.device)
        with self._guard():
            check(self.lib.sober_car_prepare(_ptr(feats), feats.stride(0), _ptr(div), S, n, _ptr(out), out.stride(0),
                                             self._stream()), "car_prepare")
        self.launches += 1
        return out

    def car_summary(self, w, delta, defect_limit=1e-5):
        """Poison ``w`` (in place) when the projector was inaccurate, then (summary int32 [S + 1], rank int32 [S])."""
        S = w.numel()
        summary = torch.empty(S + 1, dtype=torch.int32, device=self.device)
        rank = torch.empty(S, dtype=torch.int32, device=self.device)
        with self._guard():
            check(self.lib.sober_car_summary(_ptr(w), S, _ptr(delta), 0 if delta is None else delta.numel(),
                                             float(defect_limit), _ptr(summary), _ptr(rank), self._stream()), "car_summary")
        self.launches += 1
        return summary, rank

    def fetch_small(self, t):
        """Start a device-to-host copy of the small tensor ``t`` (as it is NOW in stream order) on a side stream into pinned
        memory; returns a function that waits for that copy alone and gives the values as a list.  Work enqueued on the
        current stream after this call is not waited for."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_fetch_stream", None) is None:
            self._fetch_stream = torch.cuda.Stream(self.device)
            self._fetch_buf = {}
        key = (t.dtype, t.numel())
        host = self._fetch_buf.get(key)
        if host is None:
            host = self._fetch_buf[key] = torch.empty(t.numel(), dtype=t.dtype).pin_memory()
        ready = torch.cuda.Event()
        ready.record(main)
        side = self._fetch_stream
        side.wait_event(ready)
        with torch.cuda.stream(side):
            host.copy_(t.reshape(-1), non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)

        def wait():
            done.synchronize()
            return host.tolist()
        return wait

    def apply_tail(self, at, totw, tail_at=None, tail_tw=None):
        """at[S-1] += tail_at (in place) and a copy of ``totw`` with tail_tw added to its last entry: the second count of
        the remainder (SOBER/_rchq.py:153-164) in one launch.  Without a remainder: just the copy."""
        S, Lp = at.shape
        out = torch.empty_like(totw)
        with self._guard():
            check(self.lib.sober_apply_tail(_ptr(at[S - 1]), _ptr(tail_at), Lp, _ptr(totw), _ptr(tail_tw), S, _ptr(out),
                                            self._stream()), "apply_tail")
        self.launches += 1
        return out

    # -- update + compaction ----------------------------------------------------------------------------------
    def update_compact(self, idx, mu, n_local, pos0, ES, S, wstar, totw, rank, K, tail_keep, new_pos0, n_out,
                       rec=None, d=0):
        idx_out = torch.empty(n_out, dtype=torch.int32, device=self.device)
        mu_out = torch.empty(n_out, dtype=torch.float64, device=self.device)
        ldr = 0 if rec is None else rec.stride(0)
        rec_out = None if rec is None else torch.empty((n_out, ldr), dtype=torch.float64, device=self.device)
        with self._guard():
            t0 = self._begin("update_compact")
            check(self.lib.sober_update_compact(_ptr(idx), _ptr(mu), int(n_local), int(pos0), int(ES), int(S),
                                                _ptr(wstar), _ptr(totw), _ptr(rank), int(K), int(bool(tail_keep)),
                                                int(new_pos0), _ptr(idx_out), _ptr(mu_out), _ptr(rec), _ptr(rec_out),
                                                int(ldr), int(d), self._stream()),
                  "update_compact")
            self._end("update_compact", t0, (12 + 8 * ldr) * (int(n_local) + int(n_out)))  # work = HBM bytes
        self.launches += 1
        return idx_out, mu_out, rec_out

    def update_compact_dev(self, idx, mu, n_local, pos0, ES, S, wstar, totw, rank, summary, rec=None, d=0):
        """``update_compact`` with the survivor counts read from ``summary`` on the device: no host value is needed, so
        it is enqueued right behind the Caratheodory step.  Outputs hold n_local entries; the caller slices them once
        it knows the number of survivors."""
        idx_out = torch.empty(n_local, dtype=torch.int32, device=self.device)
        mu_out = torch.empty(n_local, dtype=torch.float64, device=self.device)
        ldr = 0 if rec is None else rec.stride(0)
        rec_out = None if rec is None else torch.empty((n_local, ldr), dtype=torch.float64, device=self.device)
        with self._guard():
            t0 = self._begin("update_compact")
            check(self.lib.sober_update_compact_dev(_ptr(idx), _ptr(mu), int(n_local), int(pos0), int(ES), int(S),
                                                    _ptr(wstar), _ptr(totw), _ptr(rank), _ptr(summary), _ptr(idx_out),
                                                    _ptr(mu_out), _ptr(rec), _ptr(rec_out), int(ldr), int(d),
                                                    self._stream()), "update_compact_dev")
            self._end("update_compact", t0, (12 + 8 * ldr) * int(n_local) * 3 // 2)   # work = HBM bytes (half survive)
        self.launches += 1
        return idx_out, mu_out, rec_out

    def scatter_result(self, dst, idx, w):
        with self._guard():
            check(self.lib.sober_scatter_result(_ptr(dst), dst.numel(), _ptr(idx), _ptr(w), idx.numel(),
                                                self._stream()), "scatter_result")
        self.launches += 1

    def kmeans_assign(self, X, centroids):
        """labels (n,) int64: nearest centroid of every row of X (n x d, d <= 16); first minimum on ties."""
        n, d = X.shape
        K = centroids.shape[0]
        assert X.stride(1) == 1 and centroids.is_contiguous() and centroids.shape[1] == d
        labels = torch.empty(n, dtype=torch.int64, device=self.device)
        with self._guard():
            t0 = self._begin("kmeans_assign")
            check(self.lib.sober_kmeans_assign(_ptr(X), X.stride(0), n, d, _ptr(centroids), K, _ptr(labels),
                                               self._stream()), "kmeans_assign")
            self._end("kmeans_assign", t0, n * K)
        self.launches += 1
        return labels

    def gp_rows(self, k_rows, t_rows, alpha, mean_const, kxx_const, noise, min_var, eta, mean, var, pi):
        """Per-candidate epilogue of the GP posterior (csrc/predict.cu): mean, clamped variance and the LFI measure."""
        m, n_obs = k_rows.shape
        assert k_rows.stride(1) == 1 and (t_rows is None or t_rows.stride(1) == 1) and alpha.is_contiguous()
        with self._guard():
            t0 = self._begin("gp_rows")
            check(self.lib.sober_gp_rows(_ptr(k_rows), k_rows.stride(0), _ptr(t_rows),
                                         0 if t_rows is None else t_rows.stride(0), _ptr(alpha), m, n_obs,
                                         float(mean_const), None, float(kxx_const), float(noise), float(min_var),
                                         float(eta), _ptr(mean), _ptr(var), _ptr(pi), self._stream()), "gp_rows")
            self._end("gp_rows", t0, 16 * m * n_obs + 24 * m)                  # work = HBM bytes
        self.launches += 1

    def partition_stream(self):
        """Stream confined to all SMs but ``SOBER_B200_RESERVE_SMS`` (default 8), or None when the driver cannot
        partition the device (include/sober_b200.h: sober_partition_stream)."""
        if not hasattr(self, "_partition"):
            reserve = int(os.environ.get("SOBER_B200_RESERVE_SMS", "8"))
            self._partition = None
            if reserve > 0:
                ptr, sms = C.c_void_p(), C.c_int32()
                with self._guard():
                    _lib.check(self.lib.sober_partition_stream(reserve, C.byref(ptr), C.byref(sms)), "partition_stream")
                if ptr.value:
                    self._partition = torch.cuda.ExternalStream(ptr.value, device=self.device)
                    self.partition_sms = int(sms.value)
        return self._partition

    def fp64_probe(self, blocks, iters):
        sink = torch.zeros(1, dtype=torch.float64, device=self.device)
        with self._guard():
            check(self.lib.sober_fp64_probe(int(blocks), int(iters), _ptr(sink), self._stream()), "fp64_probe")

    def popc_probe(self, blocks, iters):
        sink = torch.zeros(1, dtype=torch.int32, device=self.device)
        with self._guard():
            check(self.lib.sober_popc_probe(int(blocks), int(iters), _ptr(sink), self._stream()), "popc_probe")
