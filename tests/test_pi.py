"""pi evaluation over the candidate set: GP posterior (SOBER/_gp.py:212-238) and the LFI
acquisition measure (SOBER/_pi.py:20-38).

CPU: the oracle restatement reproduces, bit for bit, what the reference's own ``PI`` class returned when the fixtures
were generated (tests/golden/make_golden_pi.py), and does so again live when /root/reference is present.
GPU: ``sober_b200.pi_lfi`` / ``predict`` (K1 Gram launch + DGEMM + ``sober_gp_rows``) against the fixtures."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import gp as ogp
from oracle import kernels as ok

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["pi_matern6d", "pi_rbf_ard3d", "pi_tanimoto64"]


def load(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    t = lambda k: torch.from_numpy(z[k]).to(device)
    ls = z["ls"].tolist() or None
    cov = ok.make_kernel(str(z["fam"]), ls if ls is not None else 1.0, float(z["os"])).to(device)
    model = ok.GPModel(cov, t("Xobs"), t("y"), noise=float(z["noise"]), mean_constant=float(z["const"]))
    return z, model, t


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_pi(name):
    z, model, t = load(name)
    assert ogp.current_maximum(model) == float(z["eta"])
    assert torch.equal(ogp.lfi(t("X"), model, float(z["eta"])), t("lfi"))
    mean, var = ogp.predict(t("X"), model)
    assert torch.equal(mean, t("mean")) and torch.equal(var, t("var"))


@pytest.mark.skipif(not os.path.exists("/root/reference/SOBER/_pi.py"), reason="reference tree absent")
def test_oracle_equals_live_reference_pi():
    spec = importlib.util.spec_from_file_location("mgpi", os.path.join(GOLDEN, "make_golden_pi.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ref = mg.load_reference_pi()
    try:
        z, model, t = load("pi_matern6d")
        pi = ref.PI(model, "lfi")
        assert pi.eta == ogp.current_maximum(model)
        assert torch.equal(pi(t("X")), ogp.lfi(t("X"), model, pi.eta))
        with pytest.raises(NameError):          # SOBER/_pi.py:36 uses torch without importing it
            pi(t("X"), log=True)
    finally:
        for k in list(sys.modules):
            if k == "SOBER" or k.startswith("SOBER."):
                del sys.modules[k]


def test_install_patches_pi_and_falls_back_for_unknown_models():
    import sober_b200
    called = []

    class PI:
        def __init__(self):
            self.model, self.eta = object(), 0.0          # not a describable GP

        def lfi(self, X_cand, log=False):
            called.append(log)
            return "reference"
    mod = types.ModuleType("SOBER._pi")
    mod.PI = PI
    saved = {k: sys.modules.get(k) for k in ("SOBER", "SOBER._pi")}
    sys.modules["SOBER"] = types.ModuleType("SOBER")
    sys.modules["SOBER._pi"] = mod
    try:
        assert "SOBER._pi.PI.lfi" in sober_b200.install()
        assert PI().lfi(torch.zeros(3, 2), log=True) == "reference" and called == [True]
        assert sober_b200.install() == []
        sober_b200.uninstall()
        assert not getattr(PI.lfi, "_sober_b200", False)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_describe_gp_reads_the_stand_in():
    from sober_b200._predict import describe_gp
    z, model, t = load("pi_rbf_ard3d")
    spec = describe_gp(model)
    assert spec is not None and spec.kernel.d == 3 and spec.alpha.shape == (len(z["Xobs"]),)
    assert abs(spec.noise - float(z["noise"])) < 1e-18 and spec.mean_const == float(z["const"])
    w = spec.woodbury
    k_obs = model.covar_module.forward(t("Xobs"), t("Xobs")) + float(z["noise"]) * torch.eye(len(w), dtype=torch.float64)
    assert float((w @ k_obs - torch.eye(len(w), dtype=torch.float64)).abs().max()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_pi_matches_fixture(cuda_device, name):
    import sober_b200
    z, model, t = load(name, cuda_device)
    X = t("X")
    mean, var = sober_b200.predict(X, model)
    want_mean, want_var = torch.from_numpy(z["mean"]).to(cuda_device), torch.from_numpy(z["var"]).to(cuda_device)
    scale = float(want_mean.abs().max())
    assert float((mean - want_mean).abs().max()) < 1e-10 * max(scale, 1.0) * 10
    assert float((var - want_var).abs().max()) < 1e-9 * float(want_var.abs().max())
    pi = sober_b200.pi_lfi(model, X, float(z["eta"]))
    assert float((pi - torch.from_numpy(z["lfi"]).to(cuda_device)).abs().max()) < 1e-8
    # eta from the model itself, chunked evaluation, log variant
    pi2 = sober_b200.pi_lfi(model, X)
    assert float((pi2 - pi).abs().max()) < 1e-8
    _, _, pi3 = sober_b200.gp_posterior(model, X, eta=float(z["eta"]), chunk=777)
    assert torch.equal(pi3, pi)
    logpi = sober_b200.pi_lfi(model, X, float(z["eta"]), log=True)
    assert float((logpi - (pi + torch.finfo().eps).log()).abs().max()) == 0.0
