#!/usr/bin/env python
"""Pins oracle/gspace.py against the reference's own ``ScaleMmltGP`` methods (SOBER/BASQ/_scale_mmlt.py:208-275) and
returns a reference-bound ``gspace_kernel`` for tests/golden/make_golden.py (fixture ``gspace_matern4d``).

Run in the build container only (needs /root/reference).  The unmodified ``_scale_mmlt.py`` is loaded by file path beneath
stub ``SOBER`` / ``SOBER.BASQ`` packages; its ``from .._gp import update_gp, predict, predictive_covariance`` resolves to a
stub ``_gp`` module exporting the gpytorch-free restatements (oracle/gp.py, oracle/kernels.py: the real ``_gp.py`` imports
gpytorch/botorch, absent here).  The instance is created without ``__init__`` (which fits a GP) and given a stand-in model."""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gp as ogp  # noqa: E402
from oracle import gspace as ogs  # noqa: E402
from oracle import kernels as ok  # noqa: E402

REF = os.environ.get("SOBER_REFERENCE", "/root/reference")


def load_reference_scale_mmlt(root=REF):
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    pkg = types.ModuleType("SOBER")
    pkg.__path__ = [os.path.join(root, "SOBER")]
    sys.modules["SOBER"] = pkg
    load("SOBER._settings", os.path.join(root, "SOBER", "_settings.py"))
    load("SOBER._utils", os.path.join(root, "SOBER", "_utils.py"))
    gp = types.ModuleType("SOBER._gp")
    gp.update_gp = lambda *a, **k: None
    gp.predict = ogp.predict
    gp.predictive_covariance = ok.predictive_covariance
    sys.modules["SOBER._gp"] = gp
    basq = types.ModuleType("SOBER.BASQ")
    basq.__path__ = [os.path.join(root, "SOBER", "BASQ")]
    sys.modules["SOBER.BASQ"] = basq
    return load("SOBER.BASQ._scale_mmlt", os.path.join(root, "SOBER", "BASQ", "_scale_mmlt.py"))


def reference_instance(model, mod=None):
    mod = mod or load_reference_scale_mmlt()
    inst = object.__new__(mod.ScaleMmltGP)
    inst.model = model
    inst.jitter = torch.tensor(0.0, dtype=torch.float64)      # self.tensor(0), SOBER/BASQ/_scale_mmlt.py:70
    return inst


def check_restatement():
    g = torch.Generator().manual_seed(3)
    x_obs = torch.rand(25, 4, dtype=torch.float64, generator=g)
    y_h = torch.log1p(torch.exp(-3.0 * ((x_obs - 0.5) ** 2).sum(-1)))
    model = ok.GPModel(ok.make_kernel("matern", [0.7], 0.9), x_obs, y_h, noise=1e-3, mean_constant=0.1)
    ref, mine = reference_instance(model), ogs.ScaleMmltGP(model)
    x = torch.rand(13, 4, dtype=torch.float64, generator=g)
    y2 = torch.rand(40, 4, dtype=torch.float64, generator=g)
    y3 = torch.rand(5, 8, 4, dtype=torch.float64, generator=g)
    assert torch.equal(ref.gspace_mean_predict(x), mine.gspace_mean_predict(x))
    assert torch.equal(ref.gspace_kernel(x, y2), mine.gspace_kernel(x, y2))
    assert torch.equal(ref.gspace_kernel(x, y3), mine.gspace_kernel(x, y3))
    assert torch.equal(ref.gspace_kernel(x, x), mine.gspace_kernel(x, x))
    print("oracle/gspace.py == SOBER/BASQ/_scale_mmlt.py (gspace_mean_predict, gspace_kernel 2-D / 3-D): bitwise")


if __name__ == "__main__":
    check_restatement()
