"""GPU tests of the tcgen05 Tanimoto kernel (csrc/group_bits_mma.cu): bit-packed fingerprints expanded to 0/1 bytes in
shared memory, <x, z> by `tcgen05.mma.kind::i8` with int32 accumulators in TMEM -- exact integers, so the Gram must be
BITWISE equal to the popcount kernel's (variant 4) and agree with the oracle's float Tanimoto to rounding; grouped sums
(weights, shard offset, remainder, several row splits) agree to summation-order rounding."""
import pytest
import torch

from oracle import kernels as ok

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from sober_b200._ops import CudaOps
    return CudaOps(cuda_device)


def tables(ops, X, Z):
    from sober_b200 import _lib
    from sober_b200._ops import LandmarkTable, PointSet
    xw, xp, okx = ops.pack_bits(X)
    zw, zp, okz = ops.pack_bits(Z)
    assert okx and okz
    pts = PointSet(xw, xw.stride(0), xp, 1, X.shape[0], X.shape[1])
    lm = LandmarkTable(zw, zp, _lib.TANIMOTO_BITS, 1.7, d=X.shape[1])
    return pts, lm


@pytest.mark.parametrize("d,density", [(1024, 0.05), (512, 0.3), (256, 0.1), (1000, 0.04), (300, 0.2)])
def test_mma_gram_bitwise_equals_popcount_kernel(ops, cuda_device, d, density):
    g = torch.Generator().manual_seed(d)
    n, L = 9000, 333
    X = (torch.rand(n, d, generator=g) < density).to(torch.float64).to(cuda_device)
    X[5] = 0
    Z = X[torch.randperm(n, generator=g)[:L].to(cuda_device)].clone()
    pts, lm = tables(ops, X, Z)
    out = {}
    for variant in (0, 4, 5):       # 0: landmark tile in TMEM, 5: landmark tile in shared memory, 4: popcount
        ops.variant = variant
        try:
            out[variant], _ = ops.group_accumulate(pts, lm, None, None, n, 0, 0, n)
        finally:
            ops.variant = 0
    assert torch.equal(out[0], out[4])
    assert torch.equal(out[5], out[4])
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("tanimoto", 1.0, 1.7).to(cuda_device)), mode="kernel")
    want = kern(Z, X).T
    # one Newton step on the MUFU reciprocal seed: 3e-13 measured (the kernel matrix has to match to 1e-10)
    assert float((out[0] - want).abs().max() / want.abs().max()) < 1e-12


@pytest.mark.parametrize("n_local,S,pos0", [(50_123, 1000, 777), (200_000, 1000, 0), (12_345, 200, 31), (4_100, 64, 0)])
def test_mma_group_sums_match_popcount_kernel(ops, cuda_device, n_local, S, pos0):
    g = torch.Generator().manual_seed(n_local)
    d, L = 1024, 500
    X = (torch.rand(n_local, d, generator=g) < 0.05).to(torch.float64).to(cuda_device)
    Z = X[:L].clone()
    mu = torch.rand(n_local, dtype=torch.float64, generator=g).to(cuda_device)
    mu[torch.randperm(n_local, generator=g)[:n_local // 50].to(cuda_device)] = 0.0
    idx = torch.randperm(n_local, generator=g).to(torch.int32).to(cuda_device)        # a shuffled alive-list
    pts, lm = tables(ops, X, Z)
    ES = ((pos0 + n_local) // S) * S - S                                               # a remainder of more than one row
    out = {}
    for variant in (0, 4, 5):
        ops.variant = variant
        try:
            out[variant] = ops.group_accumulate(pts, lm, idx, mu, n_local, pos0, ES, S)
        finally:
            ops.variant = 0
    at4, tw4 = out[4]
    for v in (0, 5):
        at, tw = out[v]
        assert float((at - at4).abs().max() / at4.abs().max()) < 1e-13
        assert float((tw - tw4).abs().max() / tw4.abs().max()) < 1e-13
