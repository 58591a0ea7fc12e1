"""CPU tier: the C-ABI library builds/loads and exports every symbol include/sc_b200.h declares, with the same arity as
the ctypes table; no compute call is made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(sc_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_library_exports_every_declared_symbol():
    from sparse_caption_b200 import lib
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    so = ctypes.CDLL(lib.LIB_PATH)
    decl = _header_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(so, name), f"{name} declared in include/sc_b200.h but not exported"


def test_ctypes_table_matches_header():
    from sparse_caption_b200 import lib
    decl = _header_functions()
    for name, sig in lib.SIGNATURES.items():
        assert name in decl, f"{name} bound in lib.py but not declared in the header"
        assert decl[name] == len(sig), f"{name}: header has {decl[name]} parameters, ctypes table {len(sig)}"
    for name in decl:
        assert name in lib.SIGNATURES or name in ("sc_last_error", "sc_version"), f"{name} has no ctypes binding"


def test_version_and_error_string():
    from sparse_caption_b200 import lib
    so = lib.load()
    assert so.sc_version() >= 100
    assert isinstance(so.sc_last_error(), bytes)


def test_missing_library_fails_loudly(monkeypatch):
    from sparse_caption_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libsc_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        lib.load()
