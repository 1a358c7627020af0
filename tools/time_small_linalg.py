import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
dev = torch.device("cuda")
def bench(name, fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); print("%-44s %8.3f ms" % (name, (time.perf_counter() - t) / n * 1e3))
for L, q in [(1000, 199), (2000, 999)]:
    K = torch.randn(L, L, dtype=torch.float64, device=dev); K = K @ K.T / L + torch.eye(L, dtype=torch.float64, device=dev)
    R = torch.randn(L, q, dtype=torch.float64, device=dev)
    Y = K @ R
    G = Y.T @ Y
    C = torch.linalg.cholesky(G)
    B = R.T @ K
    T = torch.triu(torch.randn(q, q, dtype=torch.float64, device=dev))
    print("L=%d q=%d" % (L, q))
    bench("K @ R (LxL @ Lxq)", lambda: K @ R)
    bench("Y^T Y", lambda: Y.T @ Y)
    bench("cholesky_ex(q)", lambda: torch.linalg.cholesky_ex(G))
    if q <= 224:
        from sober_b200._linalg import cholesky_upper, solve_right_upper
        bench("cholesky_upper (2-CTA cluster kernel)", lambda: cholesky_upper(G), n=200)
        bench("cholesky_ex(q) x200", lambda: torch.linalg.cholesky_ex(G), n=200)
        Rr = cholesky_upper(G)[0]
        bench("solve_right_upper (Lxq)", lambda: solve_right_upper(Rr, Y), n=200)
    bench("cholesky_ex(q) + int(info) sync", lambda: int(torch.linalg.cholesky_ex(G)[1]))
    bench("cholesky_ex(L)", lambda: torch.linalg.cholesky_ex(K))
    bench("solve_triangular right (Lxq)", lambda: torch.linalg.solve_triangular(C.mH, Y, upper=True, left=False))
    bench("linalg.qr (Lxq)", lambda: torch.linalg.qr(Y))
    bench("svd (q x q) default", lambda: torch.linalg.svd(T))
    bench("svd (q x q) gesvdj", lambda: torch.linalg.svd(T, driver="gesvdj"))
    bench("svd (q x q) gesvd", lambda: torch.linalg.svd(T, driver="gesvd"))
    bench("svd (q x L) default", lambda: torch.linalg.svd(B, full_matrices=False))
    bench("eigh (q x q)", lambda: torch.linalg.eigh(G))
    if L <= 1000:
        bench("eigh (L x L)", lambda: torch.linalg.eigh(K))
    D = torch.randn(2 * (q + 1), q + 1, dtype=torch.float64, device=dev)
    bench("qr complete (2(q+1) x (q+1))", lambda: torch.linalg.qr(D, mode="complete"))
