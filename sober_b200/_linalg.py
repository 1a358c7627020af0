"""Small dense helpers shared by the Nystrom range finder and the projector null space.

``cholesky_upper(g)`` returns ``(r, info)`` with ``r^T r = g``, r upper triangular.  On a CUDA float64 matrix with
q <= 224 it is the register-resident 2-CTA cluster kernel of ``csrc/chol_pair.cu``; otherwise
``torch.linalg.cholesky_ex`` (cuSOLVER potrf, ~0.13 ms at q = 200).  ``info`` is a device scalar (no host sync).

``solve_right_upper(r, y)`` solves ``X @ r = y`` for upper-triangular ``r``.  On a CUDA tensor with q <= 256 it calls
the hand-written warp-per-row kernel (``csrc/small_linalg.cu``); otherwise (CPU tensors in the host-logic tests,
larger q) it is ``torch.linalg.solve_triangular``."""
import contextlib
import ctypes as C

import torch

from . import _lib

_NO_GUARD = contextlib.nullcontext()

launches = 0   # kernels launched from this module (bench.py adds them to its gpu_launches claim)


_get_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None) or (lambda i: torch.cuda.current_stream(i).cuda_stream)
_current_device = getattr(torch._C, "_cuda_getDevice", None) or torch.cuda.current_device


def _raw_stream(device):
    index = device.index if device.index is not None else torch.cuda.current_device()
    return index, C.c_void_p(_get_raw_stream(index))


def _guard(device, index):
    return _NO_GUARD if _current_device() == index else torch.cuda.device(device)


def cholesky_upper(g):
    q = g.shape[-1]
    if g.is_cuda and g.dtype == torch.float64 and g.dim() == 2 and 0 < q <= 224:
        lib = _lib.load()
        g = g.contiguous()
        r = torch.empty_like(g)
        info = torch.empty((), dtype=torch.int32, device=g.device)
        index, stream = _raw_stream(g.device)
        with _guard(g.device, index):
            _lib.check(lib.sober_cholesky_upper(
                C.c_void_p(g.data_ptr()), g.stride(0), q, C.c_void_p(r.data_ptr()), r.stride(0),
                C.c_void_p(info.data_ptr()), stream), "cholesky_upper")
        global launches
        launches += 1
        return r, info
    chol, info = torch.linalg.cholesky_ex(g)
    return chol.mH, info


def solve_right_upper(r, y):
    q = r.shape[0]
    if y.is_cuda and y.dtype == torch.float64 and q <= 256 and y.dim() == 2:
        lib = _lib.load()
        r = r.contiguous()
        if r.data_ptr() % 16:
            r = r.clone()              # the kernel fetches R with TMA bulk copies (16-byte aligned base)
        y = y.contiguous()
        out = torch.empty_like(y)
        index, stream = _raw_stream(y.device)
        with _guard(y.device, index):
            _lib.check(lib.sober_trsm_right_upper(
                C.c_void_p(y.data_ptr()), y.stride(0), C.c_void_p(r.data_ptr()), r.stride(0), y.shape[0], q,
                C.c_void_p(out.data_ptr()), out.stride(0), stream), "trsm_right_upper")
        global launches
        launches += 1
        return out
    return torch.linalg.solve_triangular(r, y, upper=True, left=False)
