"""sober_b200.install(): every module that bound the name at import time gets the B200 function, uninstall restores.
Uses a stand-in package tree (the real SOBER package needs gpytorch/botorch, absent here); when /root/reference is
present the real ``SOBER/_rchq.py`` is loaded as well and patched in place."""
import os
import sys
import types

import pytest

import sober_b200
from sober_b200._rchq import recombination as device_fast
from sober_b200._install import _fast as fast          # what install() binds: same call, results on the caller's device


def _fake_tree(pkg):
    def ref(*a, **k):
        return "reference"
    mods = {}
    for name in (pkg, pkg + "._rchq", pkg + "._sampler", pkg + ".BASQ", pkg + ".BASQ._basq", pkg + ".FBGP",
                 pkg + ".FBGP._fully_Bayesian_gp", pkg + "._sober"):
        mods[name] = types.ModuleType(name)
    for name in (pkg + "._rchq", pkg + "._sampler", pkg + ".BASQ._basq", pkg + ".FBGP._fully_Bayesian_gp"):
        mods[name].recombination = ref           # `from ._rchq import recombination`
    return mods, ref


def test_install_rebinds_every_importer_and_uninstall_restores():
    mods, ref = _fake_tree("SOBER")
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        patched = sober_b200.install()
        assert sorted(patched) == ["SOBER.BASQ._basq", "SOBER.FBGP._fully_Bayesian_gp", "SOBER._rchq", "SOBER._sampler"]
        for name in patched:
            assert sys.modules[name].recombination is fast
        assert not hasattr(sys.modules["SOBER._sober"], "recombination")      # untouched: it calls through _sampler
        assert sober_b200.install() == []                                      # idempotent
        sober_b200.uninstall()
        for name in patched:
            assert sys.modules[name].recombination is ref
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_signature_matches_reference_source():
    """Same parameter names, order and defaults as SOBER/_rchq.py:5-14."""
    import inspect
    sig = inspect.signature(fast)
    assert str(sig) == str(inspect.signature(device_fast))
    assert list(sig.parameters) == ["pts_rec", "pts_nys", "num_pts", "kernel", "device", "dtype", "init_weights",
                                    "calc_obj"]
    assert sig.parameters["init_weights"].default is None and sig.parameters["calc_obj"].default is None
    ref_file = "/root/reference/SOBER/_rchq.py"
    if os.path.exists(ref_file):
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden",
                                                                                  "make_golden.py"))
        mg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mg)
        ref = mg.load_reference()
        assert str(inspect.signature(ref.recombination)) == str(sig)
        patched = sober_b200.install()
        assert "SOBER._rchq" in patched and ref.recombination is fast
        sober_b200.uninstall()
        assert ref.recombination is not fast
        for k in list(sys.modules):
            if k == "SOBER" or k.startswith("SOBER."):
                del sys.modules[k]
