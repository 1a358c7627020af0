"""Process-wide options of the B200 path.

The reference's own knobs are the global device/dtype of ``SOBER/_settings.py:3-22`` (which ``recombination``
obeys instead of its ``device`` / ``dtype`` arguments, ``SOBER/_rchq.py:30``).  Here the device is always the
current CUDA device and the arithmetic is float64; the extra knobs select how closely the N-independent
L x L / S x S steps imitate the reference:

``mode="parity"``  Nystrom Gram from the kernel callable itself, the reference's PSD gate (Cholesky AND bitwise
                   symmetry AND non-symmetric ``eig``; ``SOBER/_utils.py:117-157``), ``torch.svd_lowrank`` and the
                   null space from the full ``torch.linalg.svd`` (``SOBER/_rchq.py:231-234``) -- i.e. exactly the
                   reference's op sequence on this device.  Slow (``eig``) but index-identical to the reference
                   run on the same device.
``mode="fast"``    Gram from the CUDA kernel (symmetric), the gate reduced to what it does in practice for an
                   (always bitwise-asymmetric) gpytorch Gram -- ``sqrt(K*K^T)`` then Cholesky with escalating
                   jitter -- and the null space taken as the trailing columns of the orthogonal projector
                   I - Q1 Q1^T (``nullspace="projector"``; ``"qr"`` = trailing columns of the complete Householder Q is
                   also available).  Same algorithm, different (equally valid) null-space basis: weights/moments
                   invariants hold, indices are those of the oracle run with the same basis.  The projector depends on
                   the Nystrom basis U only through its row space, so the range-finder basis Q^T is used as it is,
                   without the q x q rotation by the singular vectors of Q^T K (``rotate_basis``).

Every knob can be set individually; ``SOBER_B200_MODE`` picks the preset at import.
"""
import contextlib
import os

_PRESETS = {
    "parity": dict(gram="callable", gate="reference", nullspace="svd", nystrom_qr="householder"),
    "fast": dict(gram="cuda", gate="cholesky", nullspace="projector", nystrom_qr="cholqr2"),
}


def _under_profiler():
    """Nsight Compute / CUPTI injection present: ncu stops recording (and exits with an error) at the first launch on a
    green-context stream, so the SM-partitioned overlap of the first K1 pass is switched off under a profiler."""
    return any(k in os.environ for k in ("NV_NSIGHT_INJECTION_PORT_BASE", "NV_NSIGHT_INJECTION_TRANSPORT_TYPE",
                                         "CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR",
                                         "NV_TPS_LAUNCH_TOKEN", "NSYS_PROFILING_SESSION_ID"))


class Options:
    def __init__(self, mode=None):
        self.set_mode(mode or os.environ.get("SOBER_B200_MODE", "fast"))
        self.fuse = True              # introspect Kernel objects; False forces the generic-callable path
        self.generic_chunk = 1 << 16  # candidates per Gram tile on the generic path
        self.k1_variant = 0           # 0 auto (record / bit-packed kernels when they apply), 1 force the tiled kernel
        self.fused_projection = False  # hand-written DMMA projection+barycentre kernel instead of cuBLAS DGEMM
        self.overlap = os.environ.get("SOBER_B200_OVERLAP", "0" if _under_profiler() else "1") != "0"   # first K1 pass beside the tail of the range finder (two streams); fast mode only
        self.graphs = os.environ.get("SOBER_B200_GRAPHS", "1") != "0"   # replay each Caratheodory step from a CUDA graph
        self.defer_gate = os.environ.get("SOBER_B200_DEFER_GATE", "1") != "0"   # gate test beside the range finder
        self.rotate_basis = None      # None: rotate the Nystrom basis by its singular vectors unless the null spaces come
                                      # from the projector (then only span(U) matters); True / False force it
        self.car_kernel = os.environ.get("SOBER_B200_CAR", "panel")   # "panel": blocked row-distributed cluster kernel
                                      # (csrc/car_panel.cu) for the fused-arithmetic elimination; "legacy": round-1 kernels
        self.car_panel_nb = int(os.environ.get("SOBER_B200_CAR_NB", "0"))   # 0 = automatic panel width
        self.nvtx = os.environ.get("SOBER_B200_NVTX", "0") != "0"           # an NVTX range per stage of recombination()
        # multi-GPU: from this many groups S on, the projector null space of the replicated Caratheodory step is split
        # over the ranks (_car.projector_rows_sharded, 4 collectives per call).  OFF by default: measured at C5
        # (S = 2002) on 2 and 8 GPUs the step takes the same 1.6 ms either way -- it is bound by the latency of the
        # Cholesky factorisation and of the triangular solve, which do not shrink with the row split, and the GEMMs
        # that do shrink pay for the collectives.  Kept (and tested over gloo and NCCL) for larger S.
        self.car_shard_min = int(os.environ.get("SOBER_B200_CAR_SHARD_MIN", str(1 << 30)))
        self.stats = None             # optional dict that receives per-stage timings (forces syncs)

    def set_mode(self, mode):
        if mode not in _PRESETS:
            raise ValueError("mode must be one of %s" % sorted(_PRESETS))
        self.mode = mode
        for k, v in _PRESETS[mode].items():
            setattr(self, k, v)


options = Options()


@contextlib.contextmanager
def configure(**kw):
    """Temporarily override options: ``with configure(mode="parity"): ...``"""
    saved = dict(options.__dict__)
    try:
        if "mode" in kw:
            options.set_mode(kw.pop("mode"))
        for k, v in kw.items():
            if k not in saved:
                raise AttributeError(k)
            setattr(options, k, v)
        yield options
    finally:
        options.__dict__.clear()
        options.__dict__.update(saved)
