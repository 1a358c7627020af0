"""HBM-bound streaming passes at full size: achieved GB/s vs MEASURED_PEAKS.json (copy bandwidth)."""
import json, sys, torch
sys.path.insert(0, ".")
from sober_b200._ops import CudaOps
from sober_b200._rchq import KeepMap
ops = CudaOps(); dev = ops.device
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
N, d = 10_000_000, 6
g = torch.Generator(device=dev).manual_seed(0)
X = torch.rand(N, d, dtype=torch.float64, device=dev, generator=g)
mu = torch.rand(N, dtype=torch.float64, device=dev, generator=g)
c = X[:100].mean(0).contiguous(); inv = torch.full((d,), 2.0, dtype=torch.float64, device=dev)
idx, m, R = ops.compact_nonzero(mu)
rows = []
ms = timeit(lambda: ops.compact_nonzero(mu)); rows.append(("compact_nonzero", ms, N * (8 + 8 + 12)))
ms = timeit(lambda: ops.make_records(X, c, inv, idx, m)); rows.append(("make_records (alive-list)", ms, N * (4 + 8 + 8 * d + 64)))
ms = timeit(lambda: ops.make_records(X, c, inv, None, m)); rows.append(("make_records (identity)", ms, N * (8 + 8 * d + 64)))
rec = ops.make_records(X, c, inv, idx, m).rec
S = 2000; E = N // S; ES = E * S
kept = torch.zeros(S, dtype=torch.bool); kept[::2] = True
wstar = torch.where(kept, torch.rand(S, dtype=torch.float64) + 0.1, torch.zeros(S, dtype=torch.float64)).to(dev)
totw = (torch.rand(S, dtype=torch.float64) + 0.5).to(dev)
rank = (torch.cumsum(kept.int(), 0) - kept.int()).to(torch.int32).to(dev)
km = KeepMap(kept.tolist(), S, ES); n_out = km.before(N)
ms = timeit(lambda: ops.update_compact(idx, m, N, 0, ES, S, wstar, totw, rank, km.K, km.tail_keep, 0, n_out, rec=rec, d=d))
rows.append(("update_compact(+records)", ms, (N + n_out) * (12 + 64)))
ms = timeit(lambda: ops.prepare_points(X, c, inv)); rows.append(("prepare_points", ms, N * 8 * (d + 8)))
Xb = (torch.rand(1_000_000, 1024, device=dev, generator=g) < 0.05).to(torch.float64)
ms = timeit(lambda: ops.pack_bits(Xb)); rows.append(("pack_bits 1e6 x 1024", ms, 1_000_000 * (1024 * 8 + 128 + 8)))
for name, ms, byts in rows:
    print("%-28s %8.3f ms  %8.1f GB/s  %5.1f%% of measured HBM peak (%.0f GB/s)" % (name, ms, byts / ms / 1e6, 100 * byts / ms / 1e6 / peak, peak))
