"""clock64 breakdown of the Cholesky chain (library built with -DCP_PROF, see the header of this script).
   nvcc ... -DCP_PROF -c chol_pair.cu ; link into sober_b200/csrc/libsober_prof.so ; SOBER_B200_LIB=that python tools/profile_chol.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sober_b200 import _lib
from sober_b200._linalg import cholesky_upper

lib = _lib.load()
lib.sober_cholesky_prof.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda")
for q in (200,):
    a = torch.randn(2 * q, q, dtype=torch.float64, device=dev)
    g = a.T @ a
    cholesky_upper(g)
    torch.cuda.synchronize()
    lib.sober_cholesky_prof(None, 1)
    n = 50
    for _ in range(n):
        cholesky_upper(g)
    torch.cuda.synchronize()
    out = (C.c_longlong * 8)()
    lib.sober_cholesky_prof(out, 0)
    cnt = max(out[4], 1)
    print("q=%d  steps sampled per call %.1f" % (q, cnt / n))
    print("  publish -> next owner awake : %7.1f cycles" % (out[0] / cnt))
    print("  awake -> l/lc loaded        : %7.1f cycles" % (out[1] / cnt))
    print("  update+shfl+rsqrt           : %7.1f cycles (per factored column, incl. first)" % (out[2] / (cnt + n * 7)))
    print("  scale+store+flag            : %7.1f cycles" % (out[3] / (cnt + n * 7)))
