"""Edge cases of the drop-in on the GPU: degenerate sizes, zero weights, float32 inputs, fewer landmarks than the batch,
duplicated candidates.  Invariants of SOBER/_rchq.py:5-31 only (no oracle: these inputs are chaotic or trivial)."""
import warnings

import pytest
import torch

from oracle import kernels as ok

pytestmark = pytest.mark.gpu


def _kern(dev, fam="rbf", ls=0.7):
    return ok.Kernel(ok.BareModel(ok.make_kernel(fam, [ls], 1.0).to(dev)), mode="kernel")


def _check(idx, w, mu0, b):
    assert idx.dtype == torch.int64 and len(idx) == len(w) and len(idx) <= b
    assert bool((idx[1:] > idx[:-1]).all()) and bool((w > 0).all())
    assert abs(float(w.sum()) - float(mu0.sum())) < 1e-12


def _run(*a, **k):
    import sober_b200
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return sober_b200.recombination(*a, **k)


def test_all_weights_zero(cuda_device):
    X = torch.rand(500, 3, dtype=torch.float64, device=cuda_device)
    mu = torch.zeros(500, dtype=torch.float64, device=cuda_device)
    idx, w = _run(X, X[:40], 8, _kern(cuda_device), None, None, init_weights=mu)
    assert len(idx) == 0 and len(w) == 0


def test_single_candidate_and_fewer_than_batch(cuda_device):
    X = torch.rand(5, 3, dtype=torch.float64, device=cuda_device)
    mu = torch.tensor([0.1, 0.0, 0.4, 0.3, 0.2], dtype=torch.float64, device=cuda_device)
    idx, w = _run(X, torch.rand(30, 3, dtype=torch.float64, device=cuda_device), 8, _kern(cuda_device), None, None,
                  init_weights=mu.clone())
    assert idx.tolist() == [0, 2, 3, 4] and torch.equal(w, mu[idx])
    idx, w = _run(X[:1], X, 3, _kern(cuda_device), None, None)
    assert idx.tolist() == [0] and float(w[0]) == 1.0


def test_fewer_landmarks_than_batch(cuda_device):
    g = torch.Generator().manual_seed(0)
    X = torch.rand(3000, 4, dtype=torch.float64, generator=g).to(cuda_device)
    mu = torch.rand(3000, dtype=torch.float64, generator=g).to(cuda_device)
    mu /= mu.sum()
    for mode in ("fast", "parity"):
        import sober_b200
        with sober_b200.configure(mode=mode):
            m = mu.clone()
            idx, w = _run(X, X[:10].clone(), 16, _kern(cuda_device, "matern"), None, None, init_weights=m)
        _check(idx, w, mu, 16)
        assert len(idx) <= 11            # the basis has at most L = 10 rows -> at most L + 1 points


def test_float32_inputs_and_duplicates(cuda_device):
    g = torch.Generator().manual_seed(1)
    X = torch.rand(2000, 5, generator=g)
    X[100:200] = X[0:100]                # exact duplicates
    X = X.to(cuda_device)
    mu = torch.rand(2000, generator=g).to(cuda_device)
    mu /= mu.sum()
    m = mu.clone()
    idx, w = _run(X, X[:64].clone(), 12, _kern(cuda_device), None, None, init_weights=m)
    assert w.dtype == torch.float32 and len(idx) <= 12 and bool((w > 0).all())
    assert abs(float(w.double().sum()) - float(mu.double().sum())) < 1e-5
    assert int((m != 0).sum()) == len(idx)


def test_wrong_shapes_raise(cuda_device):
    X = torch.rand(100, 3, dtype=torch.float64, device=cuda_device)
    with pytest.raises(ValueError):
        _run(X, X[:10, :2], 4, _kern(cuda_device), None, None)
    with pytest.raises(ValueError):
        _run(X, X[:10], 4, _kern(cuda_device), None, None, init_weights=torch.ones(7, device=cuda_device))
