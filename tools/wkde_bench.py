"""Weighted-KDE density at the reference's default n_kde = 4096: N = 1e6 queries in 6-D.
GPU: CUDA events around wkde_pdf (device-resident inputs, 5 calls after 2 warm-ups).  CPU: the oracle restatement of
SOBER/_wkde.py:109-145 on a 20000-query sample with all host threads, scaled to N."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import wkde as oracle_wkde
from sober_b200._wkde import wkde_pdf
from sober_b200._rchq import _ops

dev = torch.device("cuda")
d, n_kde, N = 6, 4096, 1_000_000
g = torch.Generator().manual_seed(0)
centres = torch.rand(n_kde, d, dtype=torch.float64, generator=g)
w = torch.rand(n_kde, dtype=torch.float64, generator=g); w /= w.sum()
a = torch.randn(d, d, dtype=torch.float64, generator=g)
cov = (a @ a.T / d + 0.5 * torch.eye(d, dtype=torch.float64)) * 0.01
queries = torch.rand(N, d, dtype=torch.float64, generator=g)
bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
cd, wd, covd, qd, bd = (t.to(dev) for t in (centres, w, cov, queries, bounds))
ops = _ops()
for _ in range(2):
    out = wkde_pdf(cd, wd, covd, qd, bounds=bd)
torch.cuda.synchronize()
ops.timing = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = wkde_pdf(cd, wd, covd, qd, bounds=bd)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
k1 = ops.timing_largest("group_accumulate")
ops.timing = None
pairs = N * n_kde
print("GPU  wkde_pdf: %.2f ms per call  (%.2e query-centre pairs/s); K1 launch %.2f ms = %.1f TFLOP/s algorithmic at %d flop/pair"
      % (ms, pairs / ms * 1e3, k1[0], pairs * (2 * d + 2 + 12) / (k1[0] * 1e-3) / 1e12, 2 * d + 2 + 12))
sample = 20000
t = time.perf_counter()
ref = oracle_wkde.pdf(centres, w, cov, queries[:sample], bounds=bounds)
cpu_s = time.perf_counter() - t
print("CPU  oracle (reference op sequence, %d threads): %.2f s for %d queries -> %.1f s for N = %d  (x%.0f)"
      % (torch.get_num_threads(), cpu_s, sample, cpu_s * N / sample, N, cpu_s * N / sample / (ms * 1e-3)))
err = float((out[:sample].cpu() - ref).abs().max() / ref.abs().max())
print("max relative error on the sample: %.2e" % err)
