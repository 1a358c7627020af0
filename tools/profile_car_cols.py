"""Cycle breakdown of the column-distributed cluster elimination kernel (CTA 0 / thread 0 clock64 counters)."""
import ctypes as C, sys, time, torch
sys.path.insert(0, ".")
from sober_b200 import _lib, _car
lib = _lib.load()
S, npr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 200)
g = torch.Generator().manual_seed(0)
feats = torch.randn(S, npr - 1, dtype=torch.float64, generator=g) * torch.logspace(0, -4, npr - 1, dtype=torch.float64)
design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1).cuda().contiguous()
mass0 = torch.rand(S, dtype=torch.float64, generator=g); mass0 /= mass0.sum()
prof = torch.zeros(8, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
k = S - npr
for rep in range(3):
    rows = _car.projector_rows(design)
    mass = mass0.cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.sober_car_cluster_cols_profiled(C.c_void_p(rows.data_ptr()), k, S, C.c_void_p(mass.data_ptr()), 0, None,
                                             C.c_void_p(prof.data_ptr()), st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("rc", rc, "wall ms", dt * 1e3, "kept", int((mass > 0).sum()))
names = ["search+broadcast (owner of s+1)", "-", "-", "wait for broadcast s", "mu + next column + sync", "deferred update"]
own = k / 8
cnt = [own, 1, 1, k, k, k]
for n, v, c in zip(names, prof.tolist(), cnt):
    print("%-32s %10d cycles total  %7.0f per occurrence" % (n, v, v / c))
print("total cycles / step: %.0f" % (sum(prof.tolist()) / k))
