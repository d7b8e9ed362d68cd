"""CPU: the C-ABI library builds/loads and exports every symbol include/swnerf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "swnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(swnerf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import swnerf_b200
    from swnerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export: " + s
    # the Python binding table covers the header too
    assert set(_lib.declared_symbols()) == set(syms)
    assert L.swnerf_version() == 100


def test_binding_table_matches_header_signatures():
    """Every entry of the ctypes table has as many arguments as the header's prototype, pointer arguments are bound as
    void pointers, floats as c_float and 64-bit sizes as c_int64 (a mismatch corrupts the call silently)."""
    from swnerf_b200 import _lib
    src = open(os.path.join(ROOT, "include", "swnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(swnerf_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    checked = 0
    for name, argtypes in _lib._SIG.items():
        params = [p.strip() for p in protos[name].split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p_, t in zip(params, argtypes):
            if "*" in p_:
                assert t is ctypes.c_void_p, (name, p_)
            elif p_.startswith("float"):
                assert t is ctypes.c_float, (name, p_)
            elif p_.startswith("int64_t"):
                assert t is ctypes.c_int64, (name, p_)
            elif p_.startswith("int ") or p_.startswith("unsigned"):
                assert t is ctypes.c_int, (name, p_)
        checked += 1
    assert checked == len(_lib._SIG) >= 20


def test_error_channel_without_gpu():
    from swnerf_b200 import _lib
    L = _lib.lib()
    # argument validation happens before any CUDA call, so it works on a CPU-only box
    rc = L.swnerf_searchsorted(None, None, None, 1, 1, 1, 1, 1, None)
    assert rc != 0
    assert b"null pointer" in L.swnerf_last_error()
    rc = L.swnerf_sample_pdf(ctypes.c_void_p(16), None, None, None, 1, 1, 63, 128, ctypes.c_void_p(16), None, None)
    assert rc != 0 and b"exactly one of weights / cdf" in L.swnerf_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from swnerf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libswnerf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
