"""CPU: the C-ABI library loads and exports every symbol include/hf_b200.h declares
(no compute calls: there is no GPU on the build box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from pyhf_b200 import _lib, build

    build.build()
    lib = _lib.require()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in hf_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in pyhf_b200/_lib.py"
    assert lib.hf_version() >= 100
    assert set(_lib.SIGNATURES) == set(names)


def test_struct_layouts_match_header():
    from pyhf_b200 import _lib

    # hf_plan_desc: 14 int32 + 2 double + 17 pointers ; hf_fit_opts: 2 int32, 2 double, 2 int32
    assert ctypes.sizeof(_lib.PlanDesc) == 14 * 4 + 2 * 8 + 17 * 8
    assert ctypes.sizeof(_lib.FitOpts) == 8 + 16 + 8
    o = _lib.FitOpts()
    _lib.require().hf_fit_default_opts(ctypes.byref(o))
    assert o.max_iter == 200 and o.tol_edm == 1e-11


def test_invalid_plan_is_rejected_without_a_gpu():
    """argument validation happens on the host, before any CUDA call"""
    from pyhf_b200 import _lib

    lib = _lib.require()
    d = _lib.PlanDesc()
    h = ctypes.c_void_p()
    assert lib.hf_plan_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"inconsistent" in lib.hf_last_error()
    assert lib.hf_plan_create(None, ctypes.byref(h)) == -1


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under pyhf_b200/ may import it"""
    pkg = os.path.join(ROOT, "pyhf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), fn


def test_missing_library_fails_loudly(monkeypatch):
    from pyhf_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhf_b200.so")
    with pytest.raises(_lib.NativeLibraryError):
        _lib.require()
