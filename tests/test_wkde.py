"""Weighted-KDE density: oracle pinned on fixtures produced by the unmodified reference class
(tests/golden/make_golden_wkde.py), the B200 host path checked against both."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import wkde as oracle_wkde
from sober_b200 import _install
from sober_b200._wkde import pdf_of, wkde_pdf
from _cpu_ops import TorchOps

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "wkde_*.npz")))


def load(name, device="cpu"):
    raw = np.load(os.path.join(HERE, "golden", name + ".npz"))
    t = {k: torch.from_numpy(raw[k]).to(device) for k in raw.files}
    t["bounds"] = t["bounds"] if t["bounds"].numel() else None
    return t


def rel_err(got, want):
    return float((got - want).abs().max() / want.abs().max())


def test_fixtures_present():
    assert len(FIXTURES) >= 4


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_reproduces_reference_bitwise(name):
    f = load(name)
    got = oracle_wkde.pdf(f["centres"], f["weights"], f["covariance"], f["queries"], bounds=f["bounds"])
    assert torch.equal(got, f["pdf"])
    # chunking over queries does not change a bit (each query's row is computed independently)
    got = oracle_wkde.pdf(f["centres"], f["weights"], f["covariance"], f["queries"], bounds=f["bounds"], chunk=7000)
    assert torch.equal(got, f["pdf"])


@pytest.mark.parametrize("name", FIXTURES)
def test_host_path_matches_reference(name):
    """K1 with the roles swapped (centres = weighted candidates in 4 groups, queries = landmarks), CPU test double."""
    f = load(name)
    got = wkde_pdf(f["centres"], f["weights"], f["covariance"], f["queries"], bounds=f["bounds"], ops=TorchOps())
    assert rel_err(got, f["pdf"]) < 1e-12
    assert torch.equal(got == 0, f["pdf"] == 0)            # exactly the out-of-bound rows


def test_truncation_constants_and_duck_typed_estimator():
    f = load("wkde_3d_bounded")
    g = torch.Generator().manual_seed(0)
    const = 0.5 + 0.5 * torch.rand(len(f["weights"]), dtype=torch.float64, generator=g)
    kde = types.SimpleNamespace(Xobs=f["centres"], weights=f["weights"], covariance=f["covariance"], bounds=f["bounds"],
                                compute_cdf=True, constant=const)
    want = oracle_wkde.pdf(f["centres"], f["weights"], f["covariance"], f["queries"], bounds=f["bounds"], constant=const)
    assert rel_err(pdf_of(kde, f["queries"], ops=TorchOps()), want) < 1e-12


def test_edge_shapes():
    f = load("wkde_2d_free")
    ops = TorchOps()
    assert wkde_pdf(f["centres"], f["weights"], f["covariance"], f["queries"][:0], ops=ops).shape == (0,)
    for n in (1, 2, 3, 5):                                 # centre counts that need zero-weight padding to 4 groups
        w = f["weights"][:n] / f["weights"][:n].sum()
        want = oracle_wkde.pdf(f["centres"][:n], w, f["covariance"], f["queries"])
        assert rel_err(wkde_pdf(f["centres"][:n], w, f["covariance"], f["queries"], ops=ops), want) < 1e-12
    with pytest.raises(ValueError):
        wkde_pdf(f["centres"], f["weights"][:-1], f["covariance"], f["queries"], ops=ops)


def test_install_patches_the_estimator_class():
    class WeightedKernelDensityEstimation:
        def pdf(self, X):
            return "reference"
    mod = types.ModuleType("FAKE._wkde")
    mod.WeightedKernelDensityEstimation = WeightedKernelDensityEstimation
    sys.modules["FAKE._wkde"] = mod
    try:
        patched = _install.install("FAKE")
        assert "FAKE._wkde.WeightedKernelDensityEstimation.pdf" in patched
        assert WeightedKernelDensityEstimation.pdf is _install._kde_pdf
        _install.uninstall()
        assert WeightedKernelDensityEstimation().pdf(None) == "reference"
    finally:
        del sys.modules["FAKE._wkde"]


# ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_matches_reference_fixture(name):
    dev = torch.device("cuda")
    f = load(name)
    got = wkde_pdf(f["centres"], f["weights"], f["covariance"], f["queries"], bounds=f["bounds"])
    assert got.is_cuda and got.dtype == torch.float64
    assert rel_err(got.cpu(), f["pdf"]) < 1e-10            # north-star tolerance for kernel values
    assert torch.equal(got.cpu() == 0, f["pdf"] == 0)


@pytest.mark.gpu
@pytest.mark.parametrize("d,n_kde,n_query", [(6, 4096, 200_000), (3, 1001, 50_000), (12, 257, 20_000)])
def test_gpu_matches_oracle_at_size(d, n_kde, n_query):
    """Record path (d <= 8) and indexed path (d > 8), centre counts that are not multiples of 4, vs the CPU oracle on a
    sample of the queries."""
    g = torch.Generator().manual_seed(d)
    centres = torch.rand(n_kde, d, dtype=torch.float64, generator=g)
    w = torch.rand(n_kde, dtype=torch.float64, generator=g)
    w /= w.sum()
    a = torch.randn(d, d, dtype=torch.float64, generator=g)
    cov = (a @ a.T / d + 0.5 * torch.eye(d, dtype=torch.float64)) * 0.01
    queries = torch.rand(n_query, d, dtype=torch.float64, generator=g) * 1.1 - 0.05
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
    got = wkde_pdf(centres, w, cov, queries, bounds=bounds).cpu()
    pick = torch.randperm(n_query, generator=g)[:2000]
    want = oracle_wkde.pdf(centres, w, cov, queries[pick], bounds=bounds)
    assert rel_err(got[pick], want) < 1e-10
    assert torch.equal(got[pick] == 0, want == 0)
