"""Times the cluster Cholesky kernel against cuSOLVER potrf (CUDA events, 500 back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sober_b200._linalg import cholesky_upper

dev = torch.device("cuda")
for q in (100, 160, 200, 224):
    a = torch.randn(2 * q, q, dtype=torch.float64, device=dev)
    g = a.T @ a
    for name, fn in (("cluster kernel", lambda: cholesky_upper(g)), ("cholesky_ex", lambda: torch.linalg.cholesky_ex(g))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(500):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 500 * 1e3
        print("q=%3d %-16s %7.1f us  (%5.0f ns/column)" % (q, name, us, us * 1e3 / q))
