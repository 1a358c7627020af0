"""Device time of the small dense kernels, host launch cost excluded: run under
   ncu --metrics gpu__time_duration.sum -k regex:'chol_pair|trsm_right' (durations are cold-cache/serialised)
   or standalone for CUDA-event timing of 200 back-to-back launches (includes the host pacing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sober_b200._linalg import cholesky_upper, solve_right_upper

dev = torch.device("cuda")
for m, q in ((1000, 199), (400, 200), (200, 100)):
    y = torch.randn(m, q, dtype=torch.float64, device=dev)
    g = y.T @ y
    r, _ = cholesky_upper(g)
    for name, fn in (("cholesky_upper", lambda: cholesky_upper(g)), ("solve_right_upper", lambda: solve_right_upper(r, y))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print("m=%4d q=%3d %-18s %7.1f us per call (events over 200 launches)" % (m, q, name, e0.elapsed_time(e1) / 200 * 1e3))
