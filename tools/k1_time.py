"""Time K1 passes of C2 / C5 shape for the library given by SOBER_B200_LIB (tools/k1_variants.sh), default kernel
(variant 0).  One line per library: TFLOP/s at 26 flop per pair."""
import os, sys, torch
sys.path.insert(0, ".")
from sober_b200._ops import CudaOps, LandmarkTable
ops = CudaOps()
dev = ops.device
name = os.path.basename(os.environ.get("SOBER_B200_LIB", "default"))
out = []
ref = {}
for (N, L, S) in [(1_000_000, 1000, 400), (250_000, 1000, 400), (60_000, 1000, 400), (15_000, 1000, 400), (4_000, 1000, 400), (2_000_000, 2000, 2000), (40_000, 2000, 2000)]:
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.rand(N, 6, dtype=torch.float64, device=dev, generator=g)
    mu = torch.rand(N, dtype=torch.float64, device=dev, generator=g); mu /= mu.sum()
    Z = X[:L].clone()
    c = Z.mean(0).contiguous(); inv = torch.full((6,), 5 ** 0.5 / 0.5, dtype=torch.float64, device=dev)
    v = (Z - c) * inv
    lm = LandmarkTable((-2 * v).contiguous(), (v * v).sum(-1).contiguous(), 3, 1.0)
    rec = ops.make_records(X, c, inv, None, mu).rec
    E = N // S
    for variant in (0,):
        ops.variant = variant
        try:
            for _ in range(3):
                at, tw = ops.group_accumulate(None, lm, None, None, N, 0, E * S, S, rec=rec)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                at, tw = ops.group_accumulate(None, lm, None, None, N, 0, E * S, S, rec=rec)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            out.append("%s%dk: %.3f ms %.2f TF (sum %.10e)" % ("nb " if variant == 2 else "", N // 1000, ms, N * L * 26 / ms / 1e9, float(at.sum())))
        except Exception as e:
            out.append("%s%dk: FAILED %s" % ("nb " if variant == 2 else "", N // 1000, str(e)[:60]))
print("%-26s %s" % (name, " | ".join(out)))
