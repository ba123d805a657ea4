"""CPU: the C-ABI library builds, loads and exports every symbol include/mvsmpl.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mvsmpl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mvs_[a-z_0-9]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from mvsmplfitting_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(build.OUT)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert set(_lib.EXPORTS) == set(syms)
    assert _lib.load().mvs_version() >= 100


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mvsmplfitting_b200 import _lib
    from mvsmplfitting_b200.context import FittingContext
    with pytest.raises(_lib.MvsError):
        FittingContext(0)
    h = ctypes.c_void_p()
    rc = _lib.load().mvs_create(0, ctypes.byref(h))
    assert rc == -3 and not h.value                      # MVS_ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.load().mvs_last_error(None)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mvsmplfitting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import oracle|from oracle)", src, flags=re.M), f


def test_library_was_built_from_these_sources():
    """mvs_build_id() = hash of csrc/*.cu, *.cuh and include/*.h at compile time; a stale libmvsmpl.so fails here"""
    from mvsmplfitting_b200 import _lib, build
    assert _lib.load().mvs_build_id().decode() == build.source_hash() == build.built_id()
