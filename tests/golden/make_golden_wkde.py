#!/usr/bin/env python
"""Golden fixtures for the weighted-KDE density: runs the UNMODIFIED ``SOBER/_wkde.py`` of the reference (build container
only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_wkde.py

``import SOBER`` needs gpytorch/botorch/matplotlib, which are absent: the module is loaded beneath a stub ``SOBER``
package with dummy ``matplotlib`` and ``SOBER.mvnorm`` (its numpy-1 ``Inf`` import fails on numpy 2; only the
``compute_cdf=True`` branch uses it, which the fixtures do not take)."""
import importlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SOBER_REFERENCE", "/root/reference")


def load_reference_wkde(root=REF):
    pkg = types.ModuleType("SOBER")
    pkg.__path__ = [os.path.join(root, "SOBER")]
    sys.modules["SOBER"] = pkg

    class _Any(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return type(item, (), {})
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, _Any(name))
    mv = types.ModuleType("SOBER.mvnorm")

    def _no_cdf(*a, **k):
        raise NotImplementedError("multivariate_normal_cdf is stubbed")
    mv.multivariate_normal_cdf = _no_cdf
    sys.modules["SOBER.mvnorm"] = mv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return importlib.import_module("SOBER._wkde")


def make(name, seed, n_obs, d, n_kde, n_query, bounded):
    wk = load_reference_wkde()
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_obs, d, dtype=torch.float64, generator=g)
    W = torch.rand(n_obs, dtype=torch.float64, generator=g) ** 3
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)]) if bounded else None
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kde = wk.WeightedKernelDensityEstimation(X, W, d, bounds=bounds, n_kde=n_kde)
        queries = torch.rand(n_query, d, dtype=torch.float64, generator=g) * 1.3 - 0.15     # some out of bounds
        dens = kde.pdf(queries)
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        centres=kde.Xobs.numpy(), weights=kde.weights.numpy(), covariance=kde.covariance.numpy(),
                        bounds=(bounds.numpy() if bounded else np.zeros((0,))), queries=queries.numpy(),
                        pdf=dens.numpy())
    print(name, "n_kde", kde.n_kde, "pdf range", float(dens.min()), float(dens.max()),
          "zero rows", int((dens == 0).sum()))


if __name__ == "__main__":
    make("wkde_3d_bounded", 1, 3000, 3, 256, 400, True)
    make("wkde_6d_bounded", 2, 5000, 6, 512, 300, True)
    make("wkde_2d_free", 3, 800, 2, 64, 200, False)
    make("wkde_10d_bounded", 4, 4000, 10, 300, 200, True)
