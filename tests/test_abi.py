"""CPU tests: the C-ABI library loads and exports every symbol include/b200spectral.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200spectral.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "fluidsim_b200", "libb200spectral.so")
    assert os.path.exists(path), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200spectral.h but not exported"


def test_binding_covers_every_declared_symbol():
    from fluidsim_b200 import _lib

    assert set(declared_symbols()) == set(_lib.SIGNATURES)
    assert _lib.lib.b2_version() == 100


def test_error_reporting_without_gpu_compute():
    from fluidsim_b200 import _lib

    out = ctypes.c_void_p()
    rc = _lib.lib.b2_plan_create(ctypes.byref(out), 5, 8, 8, 8, 1.0, 1.0, 1.0)
    assert rc != 0
    assert b"ndim" in _lib.lib.b2_last_error()
