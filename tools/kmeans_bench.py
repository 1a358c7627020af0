"""k-means landmark selection at BASELINE configs[1] scale: N = 1e6 candidates, K = 1000
landmarks, D = 6, 10 Lloyd iterations (the reference's (N, K, D) broadcast would need 48 GB).  CUDA events, device-resident
input; the assignment kernel's share from the ops timer; a chunked torch evaluation of the same E step for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sober_b200._kmeans import kmeans
from sober_b200._rchq import _ops

dev = torch.device("cuda")
N, K, D = 1_000_000, 1000, 6
x = torch.rand(N, D, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
ops = _ops()
kmeans(x, K, 2)
torch.cuda.synchronize()
ops.timing = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
cl, c = kmeans(x, K, 10)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
big = ops.timing_largest("kmeans_assign")
ops.timing = None
flop = 3 * D * N * K          # sub, fma (2) per coordinate pair
print("kmeans N=%d K=%d D=%d, 10 iterations: %.2f ms total; assignment kernel %.2f ms per launch = %.1f TFLOP/s (%d flop per "
      "point-centroid pair) = %.2f of 35.6 TFLOP/s FP64" % (N, K, D, ms, big[0], flop / (big[0] * 1e-3) / 1e12, 3 * D,
                                                             flop / (big[0] * 1e-3) / 1e12 / 35.6))
cent = c.clone()
torch.cuda.synchronize()
e0.record()
want = torch.cat([((x[s:s + 8192, None, :] - cent[None]) ** 2).sum(-1).argmin(1) for s in range(0, N, 8192)])
e1.record()
torch.cuda.synchronize()
print("same E step as chunked torch ops on the device: %.2f ms" % e0.elapsed_time(e1))
got = ops.kmeans_assign(x, cent)
print("labels equal to the torch evaluation: %d of %d differ" % (int((got != want).sum()), N))
