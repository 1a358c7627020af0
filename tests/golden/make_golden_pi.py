#!/usr/bin/env python
"""Golden fixtures for the pi evaluation: tests/golden/pi_*.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_pi.py

The UNMODIFIED ``SOBER/_pi.py`` is loaded by file path beneath a stub ``SOBER`` package whose ``_gp`` module exports
``predict`` = ``oracle.gp.predict`` (the real ``SOBER/_gp.py`` imports gpytorch/botorch, absent here: that arithmetic is
restated, "parity unpinned").  What the fixtures pin is the reference's own ``PI`` class: ``eta`` (:15) and ``lfi``
(:20-38) on GP stand-ins of ``oracle/kernels.py``.  (``log=True`` cannot be pinned: SOBER/_pi.py never imports ``torch``,
so its ``torch.finfo()`` at :36 raises NameError in the reference itself.)  Inputs are stored, tests never need the reference."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gp as ogp  # noqa: E402
from oracle import kernels as ok  # noqa: E402

REF = os.environ.get("SOBER_REFERENCE", "/root/reference")

CASES = {
    # name: (family, lengthscale, outputscale, d, n_obs, n_cand, noise, mean constant, binary inputs)
    "pi_matern6d": ("matern", [0.5], 1.3, 6, 120, 5000, 1e-4, 0.25, False),
    "pi_rbf_ard3d": ("rbf", [0.3, 0.6, 1.1], 0.8, 3, 60, 3000, 1e-3, -0.5, False),
    "pi_tanimoto64": ("tanimoto", None, 1.0, 64, 80, 2000, 1e-2, 0.0, True),
}


def load_reference_pi(root=REF):
    pkg = types.ModuleType("SOBER")
    pkg.__path__ = [os.path.join(root, "SOBER")]
    sys.modules["SOBER"] = pkg
    gp = types.ModuleType("SOBER._gp")
    gp.predict = ogp.predict
    sys.modules["SOBER._gp"] = gp
    spec = importlib.util.spec_from_file_location("SOBER._pi", os.path.join(root, "SOBER", "_pi.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["SOBER._pi"] = mod
    spec.loader.exec_module(mod)
    return mod


def build_model(fam, ls, os_, x_obs, y, noise, const):
    cov = ok.make_kernel(fam, ls if ls is not None else 1.0, os_)
    return ok.GPModel(cov, x_obs, y, noise=noise, mean_constant=const)


def main():
    ref = load_reference_pi()
    for name, (fam, ls, os_, d, n_obs, n_cand, noise, const, binary) in CASES.items():
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        if binary:
            x_obs = (torch.rand(n_obs, d, generator=g) < 0.3).to(torch.float64)
            x = (torch.rand(n_cand, d, generator=g) < 0.3).to(torch.float64)
        else:
            x_obs = torch.rand(n_obs, d, dtype=torch.float64, generator=g)
            x = torch.rand(n_cand, d, dtype=torch.float64, generator=g)
        y = torch.sin(3.0 * x_obs).sum(-1) + 0.1 * torch.randn(n_obs, dtype=torch.float64, generator=g)
        model = build_model(fam, ls, os_, x_obs, y, noise, const)
        pi = ref.PI(model, "lfi")
        mean, var = ogp.predict(x, model)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), fam=fam, ls=np.array(ls if ls else []), os=os_, noise=noise,
                            const=const, Xobs=x_obs.numpy(), y=y.numpy(), X=x.numpy(), eta=pi.eta,
                            lfi=pi(x).numpy(), mean=mean.numpy(), var=var.numpy())
        print(name, "eta %.6f" % pi.eta, "lfi range %.3e .. %.3e" % (float(pi(x).min()), float(pi(x).max())))


if __name__ == "__main__":
    main()
