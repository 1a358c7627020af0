"""Time the first (largest) K1 pass of C2 / C5 for the library given by SOBER_B200_LIB."""
import os, sys, torch
sys.path.insert(0, ".")
from sober_b200._ops import CudaOps, LandmarkTable
ops = CudaOps()
dev = ops.device
for (N, L, S) in [(1_000_000, 1000, 400), (2_000_000, 2000, 2000)]:
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(N, 6, dtype=torch.float64, device=dev, generator=g)
    mu = torch.rand(N, dtype=torch.float64, device=dev, generator=g); mu /= mu.sum()
    Z = X[:L].clone()
    c = Z.mean(0).contiguous(); inv = torch.full((6,), 5 ** 0.5 / 0.5, dtype=torch.float64, device=dev)
    v = (Z - c) * inv
    lm = LandmarkTable((-2 * v).contiguous(), (v * v).sum(-1).contiguous(), 3, 1.0)
    rec = ops.make_records(X, c, inv, None, mu).rec
    E = N // S
    for _ in range(3):
        at, tw = ops.group_accumulate(None, lm, None, None, N, 0, E * S, S, rec=rec)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        at, tw = ops.group_accumulate(None, lm, None, None, N, 0, E * S, S, rec=rec)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print("%-28s N=%d L=%d S=%d: %.3f ms  %.2f TFLOP/s (26 flop/pair)  checksum %.12e" % (
        os.path.basename(os.environ.get("SOBER_B200_LIB", "default")), N, L, S, ms, N * L * 26 / ms / 1e9, float(at.sum())))
