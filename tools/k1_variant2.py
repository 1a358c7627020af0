"""First thing to run when GPU time is available again: the experimental barrier-free K1 (variant 2) against the
default kernel -- bitwise-equal outputs expected (same arithmetic, same reduction order), then the timing.
Run under a short timeout (a pipeline bug would hang):   timeout 60 python tools/k1_variant2.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sober_b200._ops import CudaOps, LandmarkTable

ops = CudaOps()
dev = ops.device
for (N, L, S, pos0) in [(100_003, 1000, 400, 0), (1_000_000, 1000, 400, 0), (333_333, 530, 400, 1234), (2_000_000, 2000, 2000, 0)]:
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(N, 6, dtype=torch.float64, device=dev, generator=g)
    mu = torch.rand(N, dtype=torch.float64, device=dev, generator=g); mu /= mu.sum()
    Z = X[:L].clone()
    c = Z.mean(0).contiguous(); inv = torch.full((6,), 5 ** 0.5 / 0.5, dtype=torch.float64, device=dev)
    v = (Z - c) * inv
    lm = LandmarkTable((-2 * v).contiguous(), (v * v).sum(-1).contiguous(), 3, 1.0)
    rec = ops.make_records(X, c, inv, None, mu).rec
    ES = ((pos0 + N) // S) * S
    out = {}
    for variant in (0, 2):
        ops.variant = variant
        for _ in range(2):
            at, tw = ops.group_accumulate(None, lm, None, None, N, pos0, ES, S, rec=rec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            at, tw = ops.group_accumulate(None, lm, None, None, N, pos0, ES, S, rec=rec)
        b.record(); torch.cuda.synchronize()
        out[variant] = (at.clone(), tw.clone(), a.elapsed_time(b) / 5)
    same = torch.equal(out[0][0], out[2][0]) and torch.equal(out[0][1], out[2][1])
    print("N=%8d L=%4d S=%4d pos0=%5d: default %.3f ms, barrier-free %.3f ms (%+.1f%%), outputs bitwise equal: %s"
          % (N, L, S, pos0, out[0][2], out[2][2], 100 * (out[2][2] / out[0][2] - 1), same))
ops.variant = 0
