"""Shared helpers: load a golden fixture and rebuild its kernel object (oracle.kernels stand-ins)."""
import os

import numpy as np
import torch

from oracle import kernels as ok
from oracle import rchq as oracle_rchq

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["matern6d_rest", "matern6d_pow2", "rbf2d_branin", "rbf_ard5d", "ising24_hamming", "tanimoto256",
         "predcov_matern6d", "direct_branch", "tiny_passthrough", "objective_matern4d", "wpredcov_matern6d", "gspace_matern4d"]
LOOP_CASES = [c for c in CASES if c not in ("direct_branch", "tiny_passthrough")]


def objective(x):
    return torch.sin(3.0 * x).sum(-1) + (x ** 2).sum(-1)


class Case:
    def __init__(self, name, device="cpu"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.raw = z
        self.device = torch.device(device)
        t = lambda k: torch.from_numpy(z[k].astype(np.float64) if z[k].dtype == np.uint8 else z[k]).to(self.device)
        self.X, self.Z = t("X"), t("Z")
        self.mu = t("mu") if "mu" in z else None
        self.b = int(z["b"])
        self.fam, self.mode = str(z["fam"]), str(z["mode"])
        self.ls = z["ls"].tolist() or None
        self.os = float(z["os"])
        self.objective = objective if bool(z["objective"]) else None
        self.idx, self.w = t("idx"), t("w")
        self.U = t("U")
        self.K_raw = t("K_raw")
        self.n_car = int(z["n_car"])
        self.Xobs = t("Xobs") if "Xobs" in z else None
        self.noise = float(z["noise"]) if "noise" in z else None
        self.yobs = t("yobs") if "yobs" in z else None
        self._t = t

    def car(self, i, key):
        return self._t("car%d_%s" % (i, key))

    def kernel(self):
        cov = ok.make_kernel(self.fam, self.ls if self.ls is not None else 1.0, self.os).to(self.device)
        if self.mode == "gspace":       # BASQ's kernel: bound method of the (restated) ScaleMmltGP, SOBER/BASQ/_basq.py:55-67
            from oracle import gspace as ogs
            const = float(self.raw["const"]) if "const" in self.raw else 0.0
            return ogs.ScaleMmltGP(ok.GPModel(cov, self.Xobs, self.yobs, noise=self.noise, mean_constant=const)).gspace_kernel
        if self.mode == "kernel":
            return ok.Kernel(ok.BareModel(cov), mode="kernel")
        return ok.Kernel(ok.GPModel(cov, self.Xobs, self.yobs, noise=self.noise), mode=self.mode)


# the fast mode's null-space basis restated with LAPACK (one definition, shared with bench.py's parity block)
projector_nullspace = oracle_rchq.projector_nullspace


def fbgp_case(device="cpu"):
    """A FullyBayesianGP stand-in (oracle/fbgp.py): 12 exact GPs with different hyperparameters, distilled weights w_qd."""
    from oracle import fbgp as ofb
    g = torch.Generator().manual_seed(21)
    X = torch.rand(1500, 3, dtype=torch.float64, generator=g)
    Z = X[torch.randperm(1500, generator=g)[:40]].clone()
    mu = torch.rand(1500, dtype=torch.float64, generator=g)
    mu /= mu.sum()
    xo = torch.rand(20, 3, dtype=torch.float64, generator=g)
    yo = torch.sin(4.0 * xo).sum(-1)
    X, Z, mu, xo, yo = (t.to(device) for t in (X, Z, mu, xo, yo))
    models = [ok.GPModel(ok.make_kernel("matern", [ls], os_).to(device), xo, yo, noise=1e-3)
              for ls, os_ in ((0.3, 1.0), (0.5, 0.7), (0.8, 1.4), (1.2, 1.0), (0.4, 2.0), (0.9, 0.5),
                              (0.6, 1.1), (1.5, 0.9), (0.35, 1.3), (0.7, 0.8), (1.0, 1.6), (0.45, 0.6))]
    w_qd = torch.rand(len(models), dtype=torch.float64, generator=g)
    w_qd /= w_qd.sum()
    return X, Z, mu, models, w_qd.to(device), ofb
