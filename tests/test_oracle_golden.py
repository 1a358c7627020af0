"""The oracle (oracle/rchq.py) against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; bit-for-bit."""
import os
import warnings

import pytest
import torch

from oracle import rchq
from _cases import CASES, Case


def _run_oracle(case, trace=None):
    mu = None if case.mu is None else case.mu.clone()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(7)
        idx, w = rchq.recombination(case.X, case.Z, case.b, case.kernel(), None, None, init_weights=mu,
                                    calc_obj=case.objective, trace=trace)
    return idx, w, mu


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture_bitwise(name):
    case = Case(name)
    stages = []
    idx, w, mu = _run_oracle(case, trace=lambda s, p: stages.append((s, p)))
    assert torch.equal(idx, case.idx)
    assert torch.equal(w, case.w)
    if mu is not None:
        assert torch.equal(mu, torch.from_numpy(case.raw["mu_after"]))      # in-place mutation contract
    basis = [p for s, p in stages if s == "basis"][0]
    assert torch.equal(basis["U"], case.U)
    gram = [p for s, p in stages if s == "gram"][0]
    assert torch.equal(gram["K_raw"], case.K_raw)
    cars_in = [p for s, p in stages if s == "car_in"]
    cars_out = [p for s, p in stages if s == "car_out"]
    assert len(cars_in) == case.n_car
    for i, (cin, cout) in enumerate(zip(cars_in, cars_out)):
        assert torch.equal(cin["X"], case.car(i, "X"))
        assert torch.equal(cin["mu"], case.car(i, "mu"))
        assert torch.equal(cin["Phi"], case.car(i, "Phi"))
        assert torch.equal(cout["w"], case.car(i, "w"))
        assert torch.equal(cout["idx"], case.car(i, "idx"))


@pytest.mark.parametrize("name", CASES)
def test_fixture_invariants(name):
    """Contract of SOBER/_rchq.py:5-31 as stored: ascending idx, positive weights, mass conserved, <= b points."""
    case = Case(name)
    assert len(case.idx) <= case.b
    assert bool((case.idx[1:] > case.idx[:-1]).all())
    assert bool((case.w > 0).all())
    total = 1.0 if case.mu is None else float(case.mu.sum())
    assert abs(float(case.w.sum()) - total) < 1e-12


@pytest.mark.skipif(not os.path.isdir("/root/reference/SOBER"), reason="reference tree not present")
def test_oracle_matches_live_reference():
    """Fresh inputs (not a stored fixture) through the live reference and the oracle: identical."""
    import importlib.util, sys
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = mg.load_reference()
    from oracle import kernels as ok
    g = torch.Generator().manual_seed(123)
    X = torch.rand(2311, 4, dtype=torch.float64, generator=g)
    Z = X[:50].clone()
    mu0 = torch.rand(2311, dtype=torch.float64, generator=g)
    mu0 /= mu0.sum()
    kern = ok.Kernel(ok.BareModel(ok.make_kernel("matern", [0.7], 1.1)), mode="kernel")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(3)
        a = mu0.clone()
        idx_r, w_r = ref.recombination(X, Z, 14, kern, torch.device("cpu"), torch.float64, init_weights=a)
        torch.manual_seed(3)
        b = mu0.clone()
        idx_o, w_o = rchq.recombination(X, Z, 14, kern, None, None, init_weights=b)
    assert torch.equal(idx_r, idx_o) and torch.equal(w_r, w_o) and torch.equal(a, b)
    for k in list(sys.modules):
        if k == "SOBER" or k.startswith("SOBER."):
            del sys.modules[k]
