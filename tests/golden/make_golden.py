#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's ``SOBER/_rchq.py`` (+ ``_utils.py``, ``_settings.py``) is loaded by file path beneath a
stub ``SOBER`` package (``import SOBER`` itself needs gpytorch/botorch, which are not installed).  Its
module-level functions are wrapped -- not edited -- to record stage outputs:

* ``ker_svd_sparsify``       -> the Nystrom basis U (after the PSD gate and ``torch.svd_lowrank``)
* ``Tchernychova_Lyons_CAR`` -> inputs (X, mu), outputs (w*, idx*), and the null-space basis Phi it derived
* the kernel callable        -> the raw Nystrom Gram

The kernel callables are the gpytorch-free objects of ``oracle/kernels.py`` (gpytorch is absent: that
arithmetic is restated, see the header there); the Tanimoto similarity is additionally checked against
``SOBER/_drug_modelling.py:15-25`` imported with dummy gpytorch/botorch modules.

Each fixture stores the inputs, so tests never need the reference at run time.
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import kernels as ok  # noqa: E402

REF = os.environ.get("SOBER_REFERENCE", "/root/reference")


def load_reference(root=REF):
    """Import SOBER._settings/_utils/_rchq from source files under a stub package."""
    pkg = types.ModuleType("SOBER")
    pkg.__path__ = [os.path.join(root, "SOBER")]
    sys.modules["SOBER"] = pkg
    for name in ("_settings", "_utils", "_rchq"):
        spec = importlib.util.spec_from_file_location("SOBER." + name, os.path.join(root, "SOBER", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["SOBER." + name] = mod
        spec.loader.exec_module(mod)
    return sys.modules["SOBER._rchq"]


def load_reference_tanimoto(root=REF):
    """Import SOBER/_drug_modelling.py with dummy gpytorch/botorch so that batch_tanimoto_sim is callable."""
    class _Any(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            sub = sys.modules.get(self.__name__ + "." + item)
            return sub if sub is not None else type(item, (), {"__init__": lambda self, *a, **k: None})
    for name in ("gpytorch", "gpytorch.kernels", "gpytorch.likelihoods", "gpytorch.means",
                 "gpytorch.distributions", "botorch", "botorch.models"):
        sys.modules.setdefault(name, _Any(name))
    spec = importlib.util.spec_from_file_location("_ref_drug", os.path.join(root, "SOBER", "_drug_modelling.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("gpytorch", "gpytorch.kernels", "gpytorch.likelihoods", "gpytorch.means",
                 "gpytorch.distributions", "botorch", "botorch.models"):
        if isinstance(sys.modules.get(name), _Any):
            del sys.modules[name]
    return mod


class Recorder:
    """Wraps the reference's module functions to capture stage outputs."""

    def __init__(self, rchq):
        self.rchq = rchq
        self.stages = []
        self._orig_ker = rchq.ker_svd_sparsify
        self._orig_car = rchq.Tchernychova_Lyons_CAR
        self._orig_svd = torch.linalg.svd

    def __enter__(self):
        rec = self

        def ker(pt, s, kernel, tm):
            S, U = rec._orig_ker(pt, s, kernel, tm)
            rec.stages.append(("basis", {"U": U.clone()}))
            return S, U

        def car(X, mu, tm, DEBUG=False):
            entry = {"X": X.clone(), "mu": mu.clone()}
            captured = {}

            def svd(a, *args, **kw):
                out = rec._orig_svd(a, *args, **kw)
                captured["Vh"] = out[2]
                return out
            torch.linalg.svd = svd
            try:
                res = rec._orig_car(X, mu, tm, DEBUG)
            finally:
                torch.linalg.svd = rec._orig_svd
            n_pts, n_dim = X.shape[0], X.shape[1] + 1
            entry["Phi"] = captured["Vh"][-(n_pts - n_dim):, :].T.clone()
            entry["w"] = res[0].clone()
            entry["idx"] = res[1].clone()
            rec.stages.append(("car", entry))
            return res

        self.rchq.ker_svd_sparsify = ker
        self.rchq.Tchernychova_Lyons_CAR = car
        return self

    def __exit__(self, *exc):
        self.rchq.ker_svd_sparsify = self._orig_ker
        self.rchq.Tchernychova_Lyons_CAR = self._orig_car
        torch.linalg.svd = self._orig_svd


# ---------------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------------
def _weights(n, zeros=0, gen=None):
    w = torch.rand(n, dtype=torch.float64, generator=gen)
    if zeros:
        w[torch.randperm(n, generator=gen)[:zeros]] = 0.0
    return w / w.sum()


def case_inputs(name):
    """Returns dict(X, Z, mu or None, b, kernel spec) -- everything seeded and self-contained."""
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    f64 = torch.float64
    if name == "matern6d_rest":          # C2 shape in miniature, N not a multiple of S: remainder quirk active
        X = torch.rand(4000, 6, dtype=f64, generator=g)
        Z = X[torch.randperm(4000, generator=g)[:96]].clone()
        return dict(X=X, Z=Z, mu=_weights(4000, 37, g), b=24, fam="matern", ls=[0.5], os=1.3, mode="kernel")
    if name == "matern6d_pow2":          # N = S * 2^k : remainder-free control, feature means preserved
        X = torch.rand(48 * 64, 6, dtype=f64, generator=g)
        Z = X[torch.randperm(len(X), generator=g)[:96]].clone()
        return dict(X=X, Z=Z, mu=None, b=24, fam="matern", ls=[0.5], os=1.0, mode="kernel")
    if name == "rbf2d_branin":           # C1 in miniature: rank-deficient Gram -> PSD gate adds jitter
        X = torch.rand(3000, 2, dtype=f64, generator=g) * 5 - 2
        Z = X[torch.randperm(3000, generator=g)[:64]].clone()
        return dict(X=X, Z=Z, mu=_weights(3000, 0, g), b=16, fam="rbf", ls=[1.0], os=1.0, mode="kernel")
    if name == "rbf_ard5d":              # ARD lengthscales
        X = torch.randn(2500, 5, dtype=f64, generator=g)
        Z = X[:80].clone()
        return dict(X=X, Z=Z, mu=_weights(2500, 11, g), b=20, fam="rbf", ls=[0.9, 1.4, 0.7, 2.0, 1.1], os=0.8,
                    mode="kernel")
    if name == "ising24_hamming":        # C3 in miniature: {0,1}^24 stored as f64, RBF == exp(-Hamming/(2 l^2))
        X = (torch.rand(4096, 24, generator=g) < 0.5).to(f64)
        Z = X[torch.randperm(4096, generator=g)[:72]].clone()
        return dict(X=X, Z=Z, mu=_weights(4096, 0, g), b=18, fam="rbf", ls=[2.0], os=1.0, mode="kernel")
    if name == "tanimoto256":            # C4 in miniature: sparse fingerprints, Tanimoto
        X = (torch.rand(3000, 256, generator=g) < 0.05).to(f64)
        Z = X[torch.randperm(3000, generator=g)[:80]].clone()
        return dict(X=X, Z=Z, mu=_weights(3000, 5, g), b=20, fam="tanimoto", ls=None, os=1.7, mode="kernel")
    if name == "predcov_matern6d":       # default Sober kernel: posterior predictive covariance
        X = torch.rand(3072, 6, dtype=f64, generator=g)
        Z = X[torch.randperm(3072, generator=g)[:96]].clone()
        Xo = torch.rand(30, 6, dtype=f64, generator=g)
        return dict(X=X, Z=Z, mu=_weights(3072, 0, g), b=24, fam="matern", ls=[0.6], os=1.0,
                    mode="predictive_covariance", Xobs=Xo, noise=1e-2)
    if name == "wpredcov_matern6d":      # Kernel(model, "weighted_predictive_covariance"): m(x) cov(x, y) m(y), SOBER/_kernel.py:33-47
        X = torch.rand(3072, 6, dtype=f64, generator=g)
        Z = X[torch.randperm(3072, generator=g)[:96]].clone()
        Xo = torch.rand(30, 6, dtype=f64, generator=g)
        yo = 1.5 + torch.sin(3.0 * Xo).sum(-1) + 0.05 * torch.randn(30, dtype=f64, generator=g)
        return dict(X=X, Z=Z, mu=_weights(3072, 0, g), b=24, fam="matern", ls=[0.6], os=1.0,
                    mode="weighted_predictive_covariance", Xobs=Xo, yobs=yo, noise=1e-2)
    if name == "gspace_matern4d":        # BASQ.quadrature (SOBER/BASQ/_basq.py:55-67): kernel = ScaleMmltGP.gspace_kernel,
        X = torch.rand(2400, 4, dtype=f64, generator=g)     # uniform weights, X_nys = the first candidates
        Z = X[:64].clone()
        Xo = torch.rand(24, 4, dtype=f64, generator=g)
        yo = torch.log1p(torch.exp(-4.0 * ((Xo - 0.5) ** 2).sum(-1)))          # h-space targets log(g + 1)
        mu = torch.ones(2400, dtype=f64) / 2400
        return dict(X=X, Z=Z, mu=mu, b=16, fam="matern", ls=[0.7], os=0.9, mode="gspace", Xobs=Xo, yobs=yo,
                    noise=1e-3, const=0.05)
    if name == "direct_branch":          # n+1 < N <= 2(n+1): a single CAR on the points themselves
        X = torch.rand(40, 3, dtype=f64, generator=g)
        Z = X[:30].clone()
        return dict(X=X, Z=Z, mu=_weights(40, 3, g), b=20, fam="rbf", ls=[0.7], os=1.0, mode="kernel")
    if name == "tiny_passthrough":       # N <= n+1: nothing to do
        X = torch.rand(12, 3, dtype=f64, generator=g)
        Z = torch.rand(20, 3, dtype=f64, generator=g)
        return dict(X=X, Z=Z, mu=_weights(12, 2, g), b=16, fam="rbf", ls=[0.7], os=1.0, mode="kernel")
    if name == "objective_matern4d":     # calc_obj branch (SOBER/_rchq.py:67-69,79-106,138-150,169-196)
        X = torch.rand(2000, 4, dtype=f64, generator=g)
        Z = X[torch.randperm(2000, generator=g)[:60]].clone()
        return dict(X=X, Z=Z, mu=_weights(2000, 0, g), b=12, fam="matern", ls=[0.8], os=1.0, mode="kernel",
                    objective=True)
    raise KeyError(name)


CASES = ["matern6d_rest", "matern6d_pow2", "rbf2d_branin", "rbf_ard5d", "ising24_hamming", "tanimoto256",
         "predcov_matern6d", "direct_branch", "tiny_passthrough", "objective_matern4d", "wpredcov_matern6d", "gspace_matern4d"]


def build_kernel(spec):
    cov = ok.make_kernel(spec["fam"], spec["ls"] if spec["ls"] is not None else 1.0, spec["os"])
    if spec["mode"] == "kernel":
        return ok.Kernel(ok.BareModel(cov), mode="kernel")
    if spec["mode"] == "gspace":
        # the REFERENCE's own bound method (unmodified SOBER/BASQ/_scale_mmlt.py on a stand-in model)
        import importlib.util as _ilu
        _s = _ilu.spec_from_file_location("mggs", os.path.join(HERE, "make_golden_gspace.py"))
        _m = _ilu.module_from_spec(_s)
        _s.loader.exec_module(_m)
        model = ok.GPModel(cov, spec["Xobs"], spec["yobs"], noise=spec["noise"], mean_constant=spec["const"])
        return _m.reference_instance(model).gspace_kernel
    model = ok.GPModel(cov, spec["Xobs"], spec.get("yobs"), noise=spec["noise"])
    return ok.Kernel(model, mode=spec["mode"])


def objective(x):
    """A fixed smooth acquisition stand-in for the calc_obj branch."""
    return torch.sin(3.0 * x).sum(-1) + (x ** 2).sum(-1)


def run_reference(rchq, spec, seed=7):
    kernel = build_kernel(spec)
    mu = None if spec["mu"] is None else spec["mu"].clone()
    calc = objective if spec.get("objective") else None
    with Recorder(rchq) as rec, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.manual_seed(seed)
        idx, w = rchq.recombination(spec["X"], spec["Z"], spec["b"], kernel, torch.device("cpu"), torch.float64,
                                    init_weights=mu, calc_obj=calc)
    return idx, w, mu, rec.stages, kernel


def main():
    rchq = load_reference()
    drug = load_reference_tanimoto()
    # pin the Tanimoto restatement against the reference source
    a = (torch.rand(17, 64) < 0.2).double()
    b = (torch.rand(5, 9, 64) < 0.2).double()
    ref_t = drug.batch_tanimoto_sim(a, b).clamp_min(0)
    assert torch.equal(ref_t, ok.TanimotoKernel().forward(a, b)), "Tanimoto restatement differs from reference"

    for name in (sys.argv[1:] or CASES):       # optional: regenerate only the named fixtures
        spec = case_inputs(name)
        idx, w, mu_after, stages, kernel = run_reference(rchq, spec)
        out = {"X": spec["X"].numpy() if spec["fam"] != "tanimoto" and name != "ising24_hamming"
               else spec["X"].numpy().astype(np.uint8),
               "Z": spec["Z"].numpy() if spec["fam"] != "tanimoto" and name != "ising24_hamming"
               else spec["Z"].numpy().astype(np.uint8),
               "b": np.int64(spec["b"]), "fam": spec["fam"], "mode": spec["mode"],
               "ls": np.asarray(spec["ls"] if spec["ls"] is not None else [], dtype=np.float64),
               "os": np.float64(spec["os"]), "objective": np.bool_(bool(spec.get("objective"))),
               "idx": idx.numpy(), "w": w.numpy(),
               "K_raw": kernel(spec["Z"], spec["Z"]).numpy()}
        if spec["mu"] is not None:
            out["mu"] = spec["mu"].numpy()
            out["mu_after"] = mu_after.numpy()
        if "Xobs" in spec:
            out["Xobs"] = spec["Xobs"].numpy()
            out["noise"] = np.float64(spec["noise"])
        if "yobs" in spec:
            out["yobs"] = spec["yobs"].numpy()
        if "const" in spec:
            out["const"] = np.float64(spec["const"])
        n_car = 0
        for stage, payload in stages:
            if stage == "basis":
                out["U"] = payload["U"].numpy()
            else:
                for key, val in payload.items():
                    out["car%d_%s" % (n_car, key)] = val.numpy()
                n_car += 1
        out["n_car"] = np.int64(n_car)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-22s N=%-5d L=%-3d b=%-3d car_calls=%-2d |idx|=%-3d sum(w)=%.16f  -> %s (%.0f KB)" % (
            name, len(spec["X"]), len(spec["Z"]), spec["b"], n_car, len(idx), float(w.sum()),
            os.path.basename(path), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
