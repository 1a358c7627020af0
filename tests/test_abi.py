"""The C-ABI shared library builds, loads and exports every symbol include/sober_b200.h declares (no compute:
there is no GPU in the CPU suite)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sober_b200.h")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from sober_b200 import _lib
    return _lib


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sober_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), "missing export: " + name


def test_python_prototypes_cover_header(built):
    assert set(declared_symbols()) == set(built.PROTOTYPES)
    assert built.load().sober_abi_version() == built.ABI_VERSION == 2


def test_header_is_plain_c():
    """No torch / C++ types in the boundary: the header must compile as C."""
    src = '#include "%s"\nint main(void){return sizeof(sober_group_args) > 0 ? 0 : 1;}\n' % HEADER
    out = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-x", "c", "-", "-o", "/dev/null"], input=src.encode(),
                         capture_output=True)
    assert out.returncode == 0, out.stderr.decode()


def test_struct_layout_matches_ctypes(built):
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){printf("%%zu %%zu %%zu %%zu",' \
          'sizeof(sober_group_args), offsetof(sober_group_args, S), offsetof(sober_group_args, Zt),' \
          'offsetof(sober_group_args, variant));return 0;}\n' % HEADER
    exe = "/tmp/_sober_layout"
    subprocess.run(["gcc", "-x", "c", "-", "-o", exe], input=src.encode(), check=True)
    size, off_s, off_zt, off_var = map(int, subprocess.run([exe], capture_output=True, check=True).stdout.split())
    G = built.GroupArgs
    assert ctypes.sizeof(G) == size
    assert G.S.offset == off_s and G.Zt.offset == off_zt and G.variant.offset == off_var


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sober_b200
    from sober_b200._lib import SoberB200Error
    x = torch.rand(10, 2, dtype=torch.float64)
    with pytest.raises(SoberB200Error):
        sober_b200.recombination(x, x[:5], 3, lambda a, b: a @ b.T, None, None)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sober_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_integration_doc_struct_matches_binding(built):
    """The ctypes struct printed in INTEGRATION.md is what a host would copy: same fields, same order as _lib.GroupArgs
    (a short struct would make the library read past it), and the ABI version it asserts is the current one."""
    doc = open(os.path.join(os.path.dirname(HEADER), "..", "INTEGRATION.md")).read()
    block = doc.split("class GroupArgs(C.Structure)")[1].split("]\n")[0]
    fields = re.findall(r'\("([a-zA-Z_0-9]+)",\s*C\.c_[a-z0-9_]+\)', block)
    assert fields == [name for name, _ in built.GroupArgs._fields_]
    assert "sober_abi_version() == %d" % built.ABI_VERSION in doc
