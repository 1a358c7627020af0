"""Is a step host-bound?  wall vs CPU time of the launching thread, and the number of CUDA kernel launches per step."""
import sys, time, warnings, torch
sys.path.insert(0, ".")
import bench, sober_b200
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_rec, d, L, b, fam, ls, desc = bench.WORKLOADS[name]
dev = torch.device("cuda")
X, mu = bench.synth(name, n_rec, 100, dev); mu /= mu.sum()
Z = X[torch.randperm(n_rec, device=dev, generator=torch.Generator(device=dev).manual_seed(1))[:L]].clone()
kern = bench.make_kernel(name, dev)
def step():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        return sober_b200.recombination(X, Z, b, kern, dev, torch.float64, init_weights=mu.clone())
for _ in range(3): step()
torch.cuda.synchronize()
w0, c0 = time.perf_counter(), time.thread_time()
for _ in range(10): step()
torch.cuda.synchronize()
w1, c1 = time.perf_counter(), time.thread_time()
print("wall %.2f ms/step, CPU (this thread) %.2f ms/step" % ((w1 - w0) * 100, (c1 - c0) * 100))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
ncuda = sum(e.count for e in ev if e.device_type == torch.autograd.DeviceType.CUDA) if hasattr(torch.autograd, "DeviceType") else -1
tot_cuda = sum(getattr(e, "self_device_time_total", 0) for e in ev)
print("profiler: total device time %.2f ms" % (tot_cuda / 1e3))
rows = sorted(ev, key=lambda e: -getattr(e, "self_device_time_total", 0))[:26]
for e in rows:
    print("%-46s calls %4d  self cpu %7.2f ms  device %7.2f ms" % (e.key[:46], e.count, e.self_cpu_time_total / 1e3, getattr(e, "self_device_time_total", 0) / 1e3))
