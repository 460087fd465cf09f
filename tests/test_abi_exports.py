"""The C-ABI libraries load and export every symbol include/*.h declares; without a GPU the boundary fails loudly."""
import ctypes as C
import os
import re

import pytest

from mass_raytrace_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_every_declared_symbol():
    lib = C.CDLL(_ffi.CUDA_LIB_PATH)
    names = _declared("mrt.h", "mrt") + _declared("mrt_debug.h", "mrt")
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(_ffi.CUDA_API)  # the Python binding table covers exactly the header
    assert lib.mrt_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    lib = C.CDLL(_ffi.HOST_LIB_PATH)
    names = _declared("mrt_host.h", "mrth")
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(names) == sorted(_ffi.HOST_API)


def test_oracle_exports_every_declared_symbol(oracle):
    src = open(os.path.join(ROOT, "oracle", "oracle.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for n in sorted(set(re.findall(r"\b(orc_[a-z0-9_]+)\s*\(", src))):
        assert hasattr(oracle, n), n


def test_product_does_not_link_the_oracle():
    import subprocess

    for lib in (_ffi.CUDA_LIB_PATH, _ffi.HOST_LIB_PATH):
        out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "oracle" not in out
        syms = subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
        assert "orc_" not in syms
    pkg = os.path.join(ROOT, "mass_raytrace_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "oracle_backend" not in text and "orc_" not in text.replace("force_", ""), f


def test_no_gpu_means_loud_failure():
    """On a box without a CUDA device mrt_context_create must return an error and a message -- never a CPU fallback."""
    lib = _ffi.cuda_lib()
    h = C.c_void_p()
    rc = lib.mrt_context_create(0, None, C.byref(h))
    if rc == 0:  # running on the GPU box
        lib.mrt_context_destroy(h)
        pytest.skip("a CUDA device is present")
    assert rc < 0 and not h.value
    msg = lib.mrt_last_error(None).decode()
    assert "no CUDA device" in msg and "no CPU fallback" in msg
    from mass_raytrace_b200 import MrtError, Renderer

    with pytest.raises(MrtError):
        Renderer(0)
    assert lib.mrt_context_create(0, None, None) == -1  # MRT_E_INVALID
