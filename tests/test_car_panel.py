"""GPU tests of the panelled elimination (csrc/car_panel.cu, SOBER/_rchq.py:237-266): the blocked factorisation on an
8-CTA cluster + whole-GPU trailing updates against the oracle's step-by-step elimination on the SAME null-space
basis -- identical support, weights to 1e-9, moments preserved -- for the single-panel and the blocked paths."""
import pytest
import torch

from oracle import rchq as oracle
from _cases import projector_nullspace

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from sober_b200._ops import CudaOps
    return CudaOps(cuda_device)


def problem(S, n_prime, seed, decades=4):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(S, n_prime - 1, dtype=torch.float64, generator=g) * \
        torch.logspace(0, -decades, n_prime - 1, dtype=torch.float64)
    mass = torch.rand(S, dtype=torch.float64, generator=g)
    mass /= mass.sum()
    design = torch.cat([torch.ones(S, 1, dtype=torch.float64), feats], 1)
    return design, mass


# (S, n', nb_hint): nb_hint 0 = automatic (single panel when it fits), else forced panel width
SHAPES = [(400, 200, 0), (200, 100, 0), (33, 7, 0), (96, 41, 0), (401, 199, 0), (17, 16, 0),
          (400, 200, 64), (400, 200, 24), (200, 100, 8), (130, 61, 64),
          (1000, 500, 0), (2000, 1000, 0), (2000, 1001, 0), (1536, 600, 32), (2048, 1024, 0), (777, 390, 0)]


@pytest.mark.parametrize("S,n_prime,nb", SHAPES)
def test_car_panel_matches_oracle_elimination(ops, cuda_device, S, n_prime, nb):
    design, mass = problem(S, n_prime, 7 * S + n_prime + nb)
    phi = projector_nullspace(design)                       # S x k, LAPACK
    k = phi.shape[1]
    fits = ops.car_panel_fits(S, k)
    assert fits in (1, 2)
    got = mass.clone().to(cuda_device)
    info = ops.car_panel(phi.T.contiguous().to(cuda_device), got, nb_hint=nb, want_info=True)
    torch.cuda.synchronize()
    got = got.cpu()
    stopped, steps = info.tolist()
    want = mass.clone()
    oracle.eliminate(phi.clone(), want, oracle.Factory())
    assert stopped == 0 and steps == k
    assert int((got > 0).sum()) <= n_prime
    assert float(got.min()) >= 0.0
    assert float((design.T @ got - design.T @ mass).abs().max()) < 1e-11
    assert torch.equal(got > 0, want > 0)
    assert float((got - want).abs().max()) < 1e-9


def test_car_panel_shape_limit(ops):
    """S <= 2048 (eight rows per lane of the pivot warp); larger problems stay with the whole-GPU kernel."""
    assert ops.car_panel_fits(2048, 1024) == 2 and ops.car_panel_fits(2049, 1024) == 0
    assert ops.car_panel_fits(400, 200) == 1


def test_car_panel_early_stop_guard(ops, cuda_device):
    """No positive entry in the leading null vector -> stop (SOBER/_rchq.py:241-242), also in a later panel."""
    rows = -torch.ones((3, 8), dtype=torch.float64, device=cuda_device)
    mass = torch.full((8,), 0.125, dtype=torch.float64, device=cuda_device)
    info = ops.car_panel(rows.clone(), mass, want_info=True)
    assert info.tolist() == [1, 0]
    assert torch.equal(mass, torch.full_like(mass, 0.125))
    # blocked: the stop happens inside the second 16-wide panel
    S, k = 64, 24
    g = torch.Generator().manual_seed(3)
    phi = torch.randn(S, k, dtype=torch.float64, generator=g)
    phi[:, :20] = 0.0
    phi[torch.arange(20), torch.arange(20)] = 1.0          # steps 0..19 remove rows 0..19 and change nothing else
    phi[:, 20] = -phi[:, 20].abs() - 1.0                   # step 20 (second 16-wide panel): no positive entry
    mass0 = torch.rand(S, dtype=torch.float64, generator=g)
    want = mass0.clone()
    removed = oracle.eliminate(phi.clone(), want, oracle.Factory())
    got = mass0.clone().to(cuda_device)
    info = ops.car_panel(phi.T.contiguous().to(cuda_device), got, nb_hint=16, want_info=True)
    stopped, steps = info.tolist()
    assert float((got.cpu() - want).abs().max()) < 1e-12
    assert steps == len(removed) and stopped == int(len(removed) < k)


def test_fast_mode_car_kernels_agree(ops, cuda_device):
    """The panelled kernel and the round-1 column-distributed kernel give the same reduction through _car.caratheodory."""
    import sober_b200
    from sober_b200 import _car
    design, mass = problem(400, 200, 99)
    feats = design[:, 1:].to(cuda_device)
    out = {}
    for name in ("panel", "legacy"):
        with sober_b200.configure(car_kernel=name):
            out[name] = _car.caratheodory(ops, feats, mass.to(cuda_device), "projector").cpu()
    assert torch.equal(out["panel"] > 0, out["legacy"] > 0)
    assert float((out["panel"] - out["legacy"]).abs().max()) < 1e-10
