"""Rebinding ``recombination`` inside an imported SOBER package.

Every consumer binds the NAME at import time (``from ._rchq import recombination``: SOBER/_sampler.py:7,
SOBER/BASQ/_basq.py:2, SOBER/FBGP/_fully_Bayesian_gp.py:10), so replacing ``SOBER._rchq.recombination`` alone is
not enough: the attribute is swapped in every already-imported module that holds it.  ``_sober.py``,
``_sober_wrapper.py`` and the ``examples/`` scripts then run unmodified.
"""
import sys

from ._rchq import recombination as _device_recombination


def _fast(pts_rec, pts_nys, num_pts, kernel, device, dtype, init_weights=None, calc_obj=None):
    """``sober_b200.recombination`` for SOBER's callers: (idx, w) come back on the device of ``pts_rec`` (callers index
    ``X_cand[idx]`` right away, SOBER/_sober.py:181); the reference returns them on its one global device."""
    idx, w = _device_recombination(pts_rec, pts_nys, num_pts, kernel, device, dtype, init_weights=init_weights,
                                   calc_obj=calc_obj)
    return idx.to(pts_rec.device), w.to(pts_rec.device)


_fast.__signature_source__ = _device_recombination

_MODULES = ("SOBER._rchq", "SOBER._sampler", "SOBER.BASQ._basq", "SOBER.FBGP._fully_Bayesian_gp")
_saved = {}
_saved_pdf = {}
_saved_fn = {}


def _kmeans(x, K=10, Niter=10):
    """Drop-in for ``KMeans`` (SOBER/_weights.py:100-126): same (labels, centroids), on x's device and dtype."""
    from ._kmeans import kmeans
    cl, c = kmeans(x, K, Niter)
    return cl.to(x.device), c.to(device=x.device, dtype=x.dtype)


def _kde_pdf(self, X):
    """Drop-in for ``WeightedKernelDensityEstimation.pdf`` (SOBER/_wkde.py:109-145): same values, returned where and
    as what the estimator keeps its own tensors."""
    from ._wkde import pdf_of
    return pdf_of(self, X).to(device=self.weights.device, dtype=self.weights.dtype)


_saved_lfi = {}


def _make_lfi(original):
    def lfi(self, X_cand, log=False):
        """Drop-in for ``PI.lfi`` (SOBER/_pi.py:20-38): the B200 path when the model is an exact GP it can describe,
        the reference's own code otherwise."""
        from ._predict import describe_gp, pi_lfi
        if describe_gp(self.model) is None:
            return original(self, X_cand, log=log)
        out = pi_lfi(self.model, X_cand, self.eta, log)
        return out.to(device=X_cand.device, dtype=X_cand.dtype)
    lfi._sober_b200 = True
    return lfi


def install(package="SOBER"):
    """Swap the reference's ``recombination`` for the B200 one; returns the list of modules patched."""
    patched = []
    for name in _MODULES:
        name = name.replace("SOBER", package, 1)
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "recombination") and mod.recombination is not _fast:
            _saved[name] = mod.recombination
            mod.recombination = _fast
            patched.append(name)
    # the weighted-KDE density shares K1: patch the class method if the module is loaded
    name = package + "._wkde"
    mod = sys.modules.get(name)
    cls = getattr(mod, "WeightedKernelDensityEstimation", None) if mod is not None else None
    if cls is not None and getattr(cls, "pdf", None) is not _kde_pdf:
        _saved_pdf[name] = cls.pdf
        cls.pdf = _kde_pdf
        patched.append(name + ".WeightedKernelDensityEstimation.pdf")
    # pi evaluation over the candidates: PI.lfi -> GP posterior through K1 + the row epilogue
    name = package + "._pi"
    mod = sys.modules.get(name)
    cls = getattr(mod, "PI", None) if mod is not None else None
    if cls is not None and not getattr(cls.lfi, "_sober_b200", False):
        _saved_lfi[name] = cls.lfi
        cls.lfi = _make_lfi(cls.lfi)
        patched.append(name + ".PI.lfi")
    # k-means landmark selection: kmeans_resampling looks KMeans up in its module globals
    name = package + "._weights"
    mod = sys.modules.get(name)
    if mod is not None and hasattr(mod, "KMeans") and mod.KMeans is not _kmeans:
        _saved_fn[name] = mod.KMeans
        mod.KMeans = _kmeans
        patched.append(name + ".KMeans")
    return patched


def uninstall():
    for name, fn in list(_saved_lfi.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.PI.lfi = fn
        del _saved_lfi[name]
    for name, fn in list(_saved_fn.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.KMeans = fn
        del _saved_fn[name]
    for name, fn in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.recombination = fn
        del _saved[name]
    for name, fn in list(_saved_pdf.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.WeightedKernelDensityEstimation.pdf = fn
        del _saved_pdf[name]
