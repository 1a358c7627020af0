"""FP64 peaks of the box: vector pipe (dependent-chain DFMA) vs tensor pipe (DMMA m8n8k4)."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from sober_b200 import _lib
lib = _lib.load()
sink = torch.zeros(1, dtype=torch.float64, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(fn, blocks, iters, flops):
    for _ in range(2): fn(blocks, iters, C.c_void_p(sink.data_ptr()), st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(blocks, iters, C.c_void_p(sink.data_ptr()), st); b.record(); torch.cuda.synchronize()
    return flops / (a.elapsed_time(b) * 1e-3) / 1e12
blocks = 148 * 8
print("DFMA (vector FP64 pipe): %.2f TFLOP/s" % run(lib.sober_fp64_probe, blocks, 1 << 16, blocks * 256 * (1 << 16) * 16))
it = 1 << 12
print("DMMA m8n8k4 (FP64 tensor pipe): %.2f TFLOP/s" % run(lib.sober_dmma_probe, blocks, it, blocks * 8 * it * 8 * 512))
