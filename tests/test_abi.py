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


def test_plain_c_host_compiles_and_links_against_the_abi(tmp_path):
    """examples/host_c/step_ns3d.c drives the fused path from plain C (CUDA runtime + the header only):
    it must compile without warnings about the ABI and link against the shared library.  Running it needs
    a GPU; without one it has to fail loudly in b2_plan_create (no CPU fallback)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        import pytest

        pytest.skip("no gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "step_ns3d")
    libdir = os.path.join(ROOT, "fluidsim_b200")
    cmd = ["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "examples", "host_c", "step_ns3d.c"), "-o", exe, "-L", libdir, "-lb200spectral",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(cmd)
    import torch

    if not torch.cuda.is_available():
        r = subprocess.run([exe, "16", "1"], capture_output=True, text=True)
        assert r.returncode != 0 and "b2_plan_create" in r.stderr
