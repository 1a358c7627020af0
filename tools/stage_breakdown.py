"""Per-stage wall-clock breakdown of one recombination call (synchronising diagnostic)."""
import sys, warnings, torch
sys.path.insert(0, ".")
import bench, sober_b200
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n_rec, d, L, b, fam, ls, desc = bench.WORKLOADS[name]
dev = torch.device("cuda")
X, mu = bench.synth(name, n_rec, 100, dev); mu /= mu.sum()
Z = X[torch.randperm(n_rec, device=dev, generator=torch.Generator(device=dev).manual_seed(1))[:L]].clone()
kern = bench.make_kernel(name, dev)
for rep in range(3):
    stats = {}
    with warnings.catch_warnings(), sober_b200.configure(mode=(sys.argv[2] if len(sys.argv) > 2 else "fast"), stats=stats if rep == 2 else None):
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        sober_b200.recombination(X, Z, b, kern, dev, torch.float64, init_weights=mu.clone())
retries = stats.pop("car_retries", 0)
tot = sum(stats.values())
print(desc, "| CAR retries with the Householder basis:", retries)
for k, v in stats.items():
    print("%-18s %8.3f ms  %5.1f%%" % (k, v, 100 * v / tot))
print("%-18s %8.3f ms" % ("total (synced)", tot))
