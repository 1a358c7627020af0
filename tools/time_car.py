"""Elimination kernels side by side: panelled (csrc/car_panel.cu) vs column-distributed cluster kernel vs the
whole-GPU kernel, on projector null-space bases of the bench shapes.  us per elimination step + cycle breakdown."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sober_b200._ops import CudaOps
from sober_b200 import _car

ops = CudaOps()
dev = ops.device
shapes = [(200, 100, 0), (400, 200, 0), (400, 200, 64), (400, 200, 32), (1000, 500, 0), (1000, 500, 32),
          (2000, 1000, 0), (2000, 1000, 32), (2000, 1000, 48)]
for S, n_prime, nb in shapes:
    g = torch.Generator().manual_seed(S)
    feats = torch.randn(S, n_prime - 1, dtype=torch.float64, generator=g) * torch.logspace(0, -4, n_prime - 1, dtype=torch.float64)
    mass = torch.rand(S, dtype=torch.float64, generator=g); mass /= mass.sum()
    design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1).to(dev)
    rows = _car.projector_rows(design)
    k = rows.shape[0]
    mass = mass.to(dev)

    def timed(fn, reps=5):
        outs = None
        for _ in range(2):
            outs = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(reps):
            r, m = rows.clone(), mass.clone()
            a.record(); fn(r, m); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return min(ts), m
    def panel(r=None, m=None):
        r = rows.clone() if r is None else r; m = mass.clone() if m is None else m
        ops.car_panel(r, m, nb_hint=nb); return m
    def cols(r=None, m=None):
        r = rows.clone() if r is None else r; m = mass.clone() if m is None else m
        ops.car_cols(r, m, exact=False); return m
    def whole(r=None, m=None):
        r = rows.clone() if r is None else r; m = mass.clone() if m is None else m
        ops.car_eliminate(r, m, exact=False); return m
    t_p, m_p = timed(panel)
    line = "S=%4d k=%4d nb=%2d  panel %8.3f ms = %6.3f us/step" % (S, k, nb, t_p, 1e3 * t_p / k)
    ref = None
    if ops.car_cols_fits(S, k):
        t_c, ref = timed(cols)
        line += " | cols %8.3f ms = %6.3f us/step" % (t_c, 1e3 * t_c / k)
    if nb == 0:
        t_w, ref2 = timed(whole)
        line += " | whole-GPU %8.3f ms = %6.3f us/step" % (t_w, 1e3 * t_w / k)
        ref = ref if ref is not None else ref2
    if ref is not None:
        line += " | same support %s, max|dw| %.1e" % (bool(torch.equal(m_p > 0, ref > 0)), float((m_p - ref).abs().max()))
    print(line)
    prof = torch.zeros(8, dtype=torch.int64, device=dev)
    r, m = rows.clone(), mass.clone()
    ops.car_panel(r, m, nb_hint=nb, prof=prof)
    torch.cuda.synchronize()
    p = prof.tolist()
    print("      panel-kernel cycles/step (CTA 0, pivot warp): wait %.0f | winner+update %.0f | ratio+argmin+post %.0f | u columns %.0f | block-start test %.0f | block end: barrier+G wait %.0f, solve+update %.0f"
          % (p[0] / k, p[1] / k, p[2] / k, p[3] / k, p[4] / k, p[5] / k, p[6] / k))
