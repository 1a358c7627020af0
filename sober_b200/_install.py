"""Rebinding ``recombination`` inside an imported SOBER package.

Every consumer binds the NAME at import time (``from ._rchq import recombination``: SOBER/_sampler.py:7,
SOBER/BASQ/_basq.py:2, SOBER/FBGP/_fully_Bayesian_gp.py:10), so replacing ``SOBER._rchq.recombination`` alone is
not enough: the attribute is swapped in every already-imported module that holds it.  ``_sober.py``,
``_sober_wrapper.py`` and the ``examples/`` scripts then run unmodified.
"""
import sys

from ._rchq import recombination as _fast

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
    # k-means landmark selection: kmeans_resampling looks KMeans up in its module globals
    name = package + "._weights"
    mod = sys.modules.get(name)
    if mod is not None and hasattr(mod, "KMeans") and mod.KMeans is not _kmeans:
        _saved_fn[name] = mod.KMeans
        mod.KMeans = _kmeans
        patched.append(name + ".KMeans")
    return patched


def uninstall():
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
