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
    return patched


def uninstall():
    for name, fn in list(_saved.items()):
        mod = sys.modules.get(name)
        if mod is not None:
            mod.recombination = fn
        del _saved[name]
