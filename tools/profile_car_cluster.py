"""Cycle breakdown of the cluster CAR kernel (CTA 0 / thread 0 clock64 counters)."""
import ctypes as C, sys, time, torch
sys.path.insert(0, ".")
from sober_b200 import _lib
lib = _lib.load()
S, npr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 200)
g = torch.Generator().manual_seed(0)
feats = torch.randn(S, npr - 1, dtype=torch.float64, generator=g) * torch.logspace(0, -4, npr - 1, dtype=torch.float64)
design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1).cuda().contiguous()
mass0 = torch.rand(S, dtype=torch.float64, generator=g); mass0 /= mass0.sum()
prof = torch.zeros(12, dtype=torch.int64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for rep in range(3):
    mass = mass0.cuda()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.sober_car_cluster_profiled(C.c_void_p(design.data_ptr()), None, S, npr, C.c_void_p(mass.data_ptr()), 0, None,
                                        C.c_void_p(prof.data_ptr()), st)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("rc", rc, "wall ms", dt * 1e3, "kept", int((mass > 0).sum()))
names = ["p1 sync(piv)", "p1 dots+stage", "p1 fence+sync+push", "p1 wait", "p1 scalar+update",
         "p2 dots+stage", "p2 fence+sync+push+wait", "p2 update", "p3 sync+argmin+sync", "p3 row+fence+sync+push",
         "p3 wait", "p3 update"]
steps = [npr] * 5 + [npr] * 3 + [S - npr] * 4
for n, v, k in zip(names, prof.tolist(), steps):
    print("%-28s %10d cycles total  %7.0f per step" % (n, v, v / k))
print("sum per step: p1 %.0f  p2 %.0f  p3 %.0f cycles" % (sum(prof[:5].tolist()) / npr, sum(prof[5:8].tolist()) / npr, sum(prof[8:].tolist()) / (S - npr)))
