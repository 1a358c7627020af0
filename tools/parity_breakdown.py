"""Where parity mode's time goes at C2 (L = 1000, q = 199, S = 400): the library calls of the reference's op sequence."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda")
torch.manual_seed(0)
L, q, S = 1000, 199, 400
X = torch.rand(L, 6, dtype=torch.float64, device=dev)
d2 = torch.cdist(X, X) * (5 ** 0.5) / 0.5
K = (1 + d2 + d2 * d2 / 3) * torch.exp(-d2)
K = torch.sqrt(K * K.T)

def timed(name, fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print("%-52s %8.3f ms" % (name, (time.perf_counter() - t) / reps * 1e3))

timed("torch.linalg.cholesky(K) 1000", lambda: torch.linalg.cholesky(K))
timed("(K == K.T).all()", lambda: bool((K == K.T).all()))
timed("torch.linalg.eig(K) 1000 (non-symmetric solver)", lambda: torch.linalg.eig(K), reps=2)
timed("torch.linalg.eigvalsh(K) 1000", lambda: torch.linalg.eigvalsh(K))
timed("torch.svd_lowrank(K, q=199)", lambda: torch.svd_lowrank(K, q=q))
D = torch.randn(S, q + 1, dtype=torch.float64, device=dev)
timed("torch.linalg.svd(design.T) 200 x 400, full", lambda: torch.linalg.svd(D.T))
timed("torch.linalg.qr(design, complete) 400 x 200", lambda: torch.linalg.qr(D, mode="complete"))
from sober_b200 import _psd
timed("_psd.passes(K, 'reference') with the eigvalsh band", lambda: _psd.passes(K, "reference"))
