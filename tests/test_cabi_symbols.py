"""-m "not gpu": libdad3d.so loads on a GPU-less box and exports every symbol include/dad3d.h declares (no compute)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dad3d.h")).read()
    return sorted(set(re.findall(r"DAD3D_API[^;(]*?\b(dad3d_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "dad3d_flame_decode" in syms and "dad3d_last_error" in syms and len(syms) >= 10


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "dad_3dheads_b200", "libdad3d.so"))
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libdad3d.so does not export {s}"


def test_python_binding_covers_header():
    from dad_3dheads_b200 import _lib
    assert set(_declared_symbols()) == set(_lib.SIGNATURES)
    lib = _lib.load()
    assert lib.dad3d_version() >= 100


def test_invalid_arguments_are_reported_not_crashed():
    from dad_3dheads_b200 import _lib
    lib = _lib.load()
    rc = lib.dad3d_flame_decode(None, None, 4, 0, None, None, 256.0, 1, None, 0, None)
    assert rc == -1 and b"null handle" in lib.dad3d_last_error()
    assert lib.dad3d_flame_workspace_bytes(None, 16) == 0


def test_product_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dad_3dheads_b200 import HeadMesh, _lib
    hm = HeadMesh()
    with pytest.raises(_lib.Dad3dError):
        hm.vertices_3d(torch.zeros(1, 413))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dad_3dheads_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
