"""Small dense helpers shared by the Nystrom range finder and the projector null space.

(The q x q Cholesky factorisations stay with ``torch.linalg.cholesky_ex``: a hand-written one-CTA kernel was 3x slower
than cuSOLVER's potrf at q = 200.)

``solve_right_upper(r, y)`` solves ``X @ r = y`` for upper-triangular ``r``.  On a CUDA tensor with q <= 256 it calls
the hand-written warp-per-row kernel (``csrc/small_linalg.cu``); otherwise (CPU tensors in the host-logic tests,
larger q) it is ``torch.linalg.solve_triangular``."""
import ctypes as C

import torch

from . import _lib


def solve_right_upper(r, y):
    q = r.shape[0]
    if y.is_cuda and y.dtype == torch.float64 and q <= 256 and y.dim() == 2:
        lib = _lib.load()
        r = r.contiguous()
        y = y.contiguous()
        out = torch.empty_like(y)
        with torch.cuda.device(y.device):
            _lib.check(lib.sober_trsm_right_upper(
                C.c_void_p(y.data_ptr()), y.stride(0), C.c_void_p(r.data_ptr()), r.stride(0), y.shape[0], q,
                C.c_void_p(out.data_ptr()), out.stride(0), C.c_void_p(torch.cuda.current_stream(y.device).cuda_stream)),
                "trsm_right_upper")
        return out
    return torch.linalg.solve_triangular(r, y, upper=True, left=False)
