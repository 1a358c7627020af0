"""cProfile of the host side of recombination() at BASELINE configs[1] (where does the Python time go?)."""
import cProfile, pstats, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, sober_b200
name = "c2"
n_rec, d, L, b, fam, ls, desc = bench.WORKLOADS[name]
dev = torch.device("cuda")
X, mu = bench.synth(name, n_rec, 100, dev); mu /= mu.sum()
Z = X[torch.randperm(n_rec, device=dev, generator=torch.Generator(device=dev).manual_seed(1))[:L]].clone()
kern = bench.make_kernel(name, dev)
warnings.simplefilter("ignore")
def step():
    sober_b200.recombination(X, Z, b, kern, dev, torch.float64, init_weights=mu.clone())
for _ in range(5):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
