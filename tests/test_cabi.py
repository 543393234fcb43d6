"""The C-ABI library loads and exports every symbol include/*.h declares.  No compute calls:
this runs without a GPU; with no device S_make must fail loudly (NULL), never fall back."""
import ctypes
import os
import re

import pytest

import libstb_b200

INCLUDE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")


def _declared_symbols():
    names = set()
    for fn in sorted(os.listdir(INCLUDE)):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(INCLUDE, fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src):
            name = m.group(1)
            if name not in ("defined", "sizeof", "void"):
                names.add(name)
    return sorted(names)


def test_library_loads():
    L = libstb_b200.lib()
    assert L.stb_device_count() >= 0


@pytest.mark.parametrize("name", _declared_symbols())
def test_symbol_exported(name):
    L = ctypes.CDLL(libstb_b200.LIB_PATH)
    assert hasattr(L, name), f"{name} declared in include/ but not exported by libstb_b200.so"


def test_no_cpu_fallback_without_device():
    L = libstb_b200.lib()
    if L.stb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    sp = L.S_make(100, 20, 100, 20, 0.5, libstb_b200.S_STABLE | libstb_b200.S_UVTABLE)
    assert not sp, "S_make must return NULL when no CUDA device is usable"
    assert L.stb_last_error()


def test_flags_rejected_like_reference():
    """S_make with neither S_STABLE nor S_UVTABLE returns NULL (lib/stable.c:131-132)."""
    L = libstb_b200.lib()
    assert not L.S_make(100, 20, 100, 20, 0.5, libstb_b200.S_FLOAT)
