"""Does the first K1 pass really run beside the tail of the range finder?  Prints event times after the fork."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sober_b200
from sober_b200 import _nystrom
from oracle import kernels as ok

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
N, L, b = 1_000_000, 1000, 200
X = torch.rand(N, 6, dtype=torch.float64, device=dev, generator=g)
Z = X[torch.randperm(N, device=dev, generator=g)[:L]].clone()
kern = ok.Kernel(ok.BareModel(ok.make_kernel("matern", [0.5], 1.0).to(dev)), mode="kernel")
warnings.simplefilter("ignore")
from sober_b200._rchq import _ops
print("partition stream:", _ops().partition_stream(), getattr(_ops(), "partition_sms", None))
for ov in (False, True):
    with sober_b200.configure(mode="fast", overlap=ov):
        _nystrom.SideStream.debug = None
        for _ in range(3):
            sober_b200.recombination(X, Z, b, kern, None, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            sober_b200.recombination(X, Z, b, kern, None, None)
        e1.record()
        torch.cuda.synchronize()
        print("overlap=%s  %.2f ms per call" % (ov, e0.elapsed_time(e1) / 5))
        if ov:
            _nystrom.SideStream.debug = []
            for _ in range(3):
                sober_b200.recombination(X, Z, b, kern, None, None)
            for d in _nystrom.SideStream.debug:
                print("   after fork: bulk work ends %.2f ms, body ends %.2f ms" % (d["bulk_end"], d["body_end"]))
            _nystrom.SideStream.debug = None
