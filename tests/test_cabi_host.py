"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports what include/cds.h declares."""
import ctypes
import os
import re

import pytest

from cleandiffuser_b200.engine import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    return ctypes.CDLL(cabi.lib_path())


def test_header_symbols_are_exported(built):
    header = open(os.path.join(ROOT, "include", "cds.h")).read()
    declared = set(re.findall(r"\b(cds_[a-z_]+)\s*\(", header))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    for sym in declared:
        assert hasattr(built, sym), sym


def test_struct_mirror_matches_library(built):
    assert built.cds_version() == cabi.ABI_VERSION
    assert built.cds_op_size() == ctypes.sizeof(cabi.Op)


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "_LIB_PATH", "/nonexistent/libcds.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        cabi.load()
