"""Wall-clock breakdown of the Nystrom range finder (fast mode) at C2 sizes: each step timed with a device sync
around it (so launch-bound steps show their host cost), plus the un-synced total."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sober_b200 import _nystrom, _psd
from sober_b200._linalg import cholesky_upper, solve_right_upper

dev = torch.device("cuda")
torch.manual_seed(0)
L, q = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 199)
X = torch.rand(L, 6, dtype=torch.float64, device=dev)
d2 = torch.cdist(X, X) * (5 ** 0.5) / 0.5
K = (1 + d2 + d2 * d2 / 3) * torch.exp(-d2)
K = 0.5 * (K + K.T)

def timed(name, fn, reps=20):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    print("%-46s %7.3f ms" % (name, (time.perf_counter() - t) / reps * 1e3))
    return out

warnings.simplefilter("ignore")
G = timed("psd gate (cholesky 1000 + checks)", lambda: _psd.repair(K, "cholesky", assume_asymmetric=True))
probe = torch.randn(L, q, dtype=torch.float64, device=dev)
Y = timed("K @ probe", lambda: G @ probe)
Q = timed("cholqr 1 pass", lambda: _nystrom._orthonormal_basis(Y, "cholqr2", passes=1))
Q = timed("cholqr 2 passes", lambda: _nystrom._orthonormal_basis(Y, "cholqr2", passes=2))
timed("  gram GEMM", lambda: Y.mH @ Y)
g = Y.mH @ Y
timed("  cholesky_upper", lambda: cholesky_upper(g))
r = cholesky_upper(g)[0]
timed("  cholesky info sync", lambda: int(cholesky_upper(g)[1]))
timed("  solve_right_upper", lambda: solve_right_upper(r, Y))
B = timed("small = Q^T K", lambda: Q.mH @ G)
timed("eigh(B B^T)", lambda: torch.linalg.eigh(B @ B.mH))
timed("_left_singular_vectors (eigh + sweeps)", lambda: _nystrom._left_singular_vectors(B))
timed("lowrank_basis total", lambda: _nystrom.lowrank_basis(G, q, qr="cholqr2"))
