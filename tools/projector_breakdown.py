"""Where the projector null space (sober_b200/_car.py::projector_rows) spends its time at the bench shapes."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sober_b200 import _car
from sober_b200._linalg import cholesky_upper, solve_right_upper
dev = torch.device("cuda")

def timed(name, fn, reps=20):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print("   %-44s %8.3f ms" % (name, (time.perf_counter() - t) / reps * 1e3))
    return out

for S, n_prime in [(400, 200), (2000, 1000)]:
    print("S = %d, n' = %d" % (S, n_prime))
    g = torch.Generator().manual_seed(S)
    feats = torch.randn(S, n_prime - 1, dtype=torch.float64, generator=g) * torch.logspace(0, -4, n_prime - 1, dtype=torch.float64)
    design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1).to(dev)
    dim = n_prime
    scaled = timed("column scaling", lambda: design / design.norm(dim=0, keepdim=True).clamp_min(1e-300))
    gram = timed("gram A^T A", lambda: scaled.mH @ scaled)
    r = timed("cholesky_upper", lambda: cholesky_upper(gram))[0]
    qt = timed("solve_right_upper (Q = A R^-1)", lambda: solve_right_upper(r, scaled))
    delta = timed("delta = Q^T Q", lambda: qt.mH @ qt)
    def neumann():
        d = delta.clone(); d.diagonal().sub_(1.0); inv = d @ d - d; inv.diagonal().add_(1.0); return inv
    inv = timed("Neumann I - D + D^2", neumann)
    timed("rows = -(Q2 inv) Q^T + I", lambda: -((qt[dim:, :] @ inv) @ qt.mH))
    timed("matrix_norm(delta)", lambda: torch.linalg.matrix_norm(delta))
    timed("projector_rows total", lambda: _car.projector_rows(design, with_defect=True))
