"""The C-ABI library: builds, loads without a GPU, exports every symbol include/graphslim_b200.h declares with the
declared number of arguments, and refuses to run on CPU tensors (no fallback path)."""
import os
import re

import pytest
import torch

from graphslim_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "graphslim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|void|const char\*|gs_sampler\*|gs_sample_job\*|gs_dsampler\*)\s+\*?(gs_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        out[name] = n
    return out


def test_library_builds_and_loads_without_gpu():
    build.build()
    lib = _lib.load()
    assert lib.gs_version() >= 100


def test_every_declared_symbol_is_exported_with_matching_arity():
    lib = _lib.load()
    decl = declared_functions()
    assert len(decl) >= 35
    for name, n_params in decl.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == n_params, f"{name}: header has {n_params} params"
    for name in _lib.SIGNATURES:
        assert name in decl, f"{name} bound in _lib.py but missing from the header"


def test_no_cpu_fallback():
    from graphslim_b200.ops import CudaOps
    with pytest.raises(_lib.GraphSlimLibraryError):
        CudaOps("cpu")


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgraphslim_b200.so")
    with pytest.raises(_lib.GraphSlimLibraryError):
        _lib.load()
