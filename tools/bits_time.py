"""K1 on bit-packed 1024-bit fingerprints (C4 shape: L = 1000 landmarks, S = 1000 groups): the tcgen05 kernel
(csrc/group_bits_mma.cu; variant 0 = landmark tile in TMEM, variant 5 = in shared memory) against the popcount kernel
(variant 4)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sober_b200 import _lib
from sober_b200._ops import CudaOps, LandmarkTable, PointSet
ops = CudaOps()
dev = ops.device
for (N, L, S, d) in [(1_000_000, 1000, 1000, 1024), (250_000, 1000, 1000, 1024), (1_000_000, 500, 400, 512)]:
    g = torch.Generator(device=dev).manual_seed(0)
    X = (torch.rand(N, d, device=dev, generator=g) < 0.05).to(torch.float64)
    mu = torch.rand(N, dtype=torch.float64, device=dev, generator=g); mu /= mu.sum()
    xw, xp, _ = ops.pack_bits(X)
    zw, zp, _ = ops.pack_bits(X[:L].clone())
    del X
    pts = PointSet(xw, xw.stride(0), xp, 1, N, d)
    lm = LandmarkTable(zw, zp, _lib.TANIMOTO_BITS, 1.0, d=d)
    E = N // S
    res = {}
    for variant in (0, 5, 4):
        ops.variant = variant
        for _ in range(2):
            at, tw = ops.group_accumulate(pts, lm, None, mu, N, 0, E * S, S)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            at, tw = ops.group_accumulate(pts, lm, None, mu, N, 0, E * S, S)
        b.record(); torch.cuda.synchronize()
        res[variant] = (a.elapsed_time(b) / 3, at.clone())
    ops.variant = 0
    ms0, ms5, ms4 = res[0][0], res[5][0], res[4][0]
    pairs = N * L
    print("N=%8d L=%4d S=%4d d=%4d: tcgen05/TMEM-A %.3f ms (%.1f G pairs/s, %.0f TOP/s int8) | tcgen05/smem-A %.3f ms | popcount %.3f ms (%.1f G pairs/s) | x%.1f | max rel diff %.1e %.1e"
          % (N, L, S, d, ms0, pairs / ms0 / 1e6, 2 * pairs * d / ms0 / 1e9, ms5, ms4, pairs / ms4 / 1e6, ms4 / ms0,
             float((res[0][1] - res[4][1]).abs().max() / res[4][1].abs().max()),
             float((res[5][1] - res[4][1]).abs().max() / res[4][1].abs().max())))
