"""k-means landmark selection (SOBER/_weights.py:100-126): oracle pinned on fixtures of the
unmodified reference function, the B200 path (hand-written assignment kernel) checked against both."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import kmeans as oracle_kmeans
from sober_b200 import _install
from sober_b200._kmeans import kmeans
from _cpu_ops import TorchOps

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "kmeans_*.npz")))


def load(name):
    raw = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return (torch.from_numpy(raw["x"]), int(raw["K"]), int(raw["Niter"]), torch.from_numpy(raw["cl"]),
            torch.from_numpy(raw["c"]))


def same_centroids(got, want, tol):
    nan_g, nan_w = torch.isnan(got), torch.isnan(want)
    return torch.equal(nan_g, nan_w) and float((got[~nan_g] - want[~nan_w]).abs().max()) <= tol


def test_fixtures_present():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_reference_bitwise(name):
    x, K, niter, cl, c = load(name)
    got_cl, got_c = oracle_kmeans.kmeans(x.clone(), K, niter, chunk=777)
    assert torch.equal(got_cl, cl)
    assert torch.equal(torch.isnan(got_c), torch.isnan(c)) and torch.equal(got_c[~torch.isnan(c)], c[~torch.isnan(c)])


@pytest.mark.parametrize("name", FIXTURES)
def test_host_path_matches_reference(name):
    x, K, niter, cl, c = load(name)
    got_cl, got_c = kmeans(x, K, niter, ops=TorchOps())
    assert torch.equal(got_cl, cl)
    assert same_centroids(got_c, c, 1e-12)


def test_fewer_points_than_clusters_is_an_error():
    with pytest.raises(ValueError):
        kmeans(torch.rand(5, 3, dtype=torch.float64), K=8, ops=TorchOps())


def test_install_patches_the_module_function():
    mod = types.ModuleType("FAKE._weights")
    mod.KMeans = lambda x, K=10, Niter=10: "reference"
    sys.modules["FAKE._weights"] = mod
    try:
        assert "FAKE._weights.KMeans" in _install.install("FAKE")
        assert mod.KMeans is _install._kmeans
        _install.uninstall()
        assert mod.KMeans(None) == "reference"
    finally:
        del sys.modules["FAKE._weights"]


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_matches_reference_fixture(name):
    """Labels identical (first-minimum and first-NaN rules included), centroids to 1e-10 (atomics reorder the sums)."""
    x, K, niter, cl, c = load(name)
    got_cl, got_c = kmeans(x.cuda(), K, niter)
    assert got_cl.is_cuda and got_cl.dtype == torch.int64
    assert torch.equal(got_cl.cpu(), cl)
    assert same_centroids(got_c.cpu(), c, 1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k", [(200_000, 6, 1000), (50_001, 3, 1025), (30_000, 16, 300), (1000, 1, 7)])
def test_gpu_assignment_kernel_vs_torch(n, d, k):
    """sober_kmeans_assign alone (several centroid chunks, padded leading dimension, d up to 16) against a chunked
    float64 torch evaluation of the same rule on the device."""
    from sober_b200._rchq import _ops
    ops = _ops()
    g = torch.Generator().manual_seed(n + d)
    wide = torch.rand(n, d + 3, dtype=torch.float64, generator=g).cuda()
    x = wide[:, :d]                                        # ldx = d + 3
    c = torch.rand(k, d, dtype=torch.float64, generator=g).cuda()
    if k > 5:
        c[5] = c[2]                                        # an exact tie: the first index must win
    got = ops.kmeans_assign(x, c)
    want = torch.cat([((x[s:s + 4096, None, :] - c[None]) ** 2).sum(-1).argmin(1) for s in range(0, n, 4096)])
    differ = got != want
    if bool(differ.any()):                                 # only near-ties (fma vs mul+add rounding) may differ
        xs, a, b = x[differ], c[got[differ]], c[want[differ]]
        da, db = ((xs - a) ** 2).sum(-1), ((xs - b) ** 2).sum(-1)
        assert float(((da - db).abs() / db).max()) < 1e-14 and int(differ.sum()) < 5
    assert not bool((got == 5).any()) or k <= 5
