"""The C-ABI library loads and exports every entry point include/lumen_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

import lumenrenderer_b200 as lr
from lumenrenderer_b200 import api
from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lumen_b200.h")).read()
    return sorted(set(re.findall(r"LB_API\s+[\w\s\*]+?\b(lb_\w+)\s*\(", text)))


def test_header_declares_the_bound_entry_points():
    syms = header_symbols()
    assert len(syms) >= 40
    bound = set("lb_" + s for s in lr.C_ABI_SYMBOLS) | set("lb_" + s for s in api.HOST_ONLY_SYMBOLS)
    assert bound == set(syms), f"ctypes binding and header disagree: {bound ^ set(syms)}"


def test_library_exports_every_declared_symbol():
    assert os.path.exists(lr.LIB_PATH), "liblumen_b200.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(lr.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.lb_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.lb_version()


def test_oracle_mirrors_the_abi(oracle):
    """Every renderer entry point has an oracle counterpart; the host-only asset ingest (lb_gltf_*) has none — its checker is the
    numpy restatement in tests/gltf_tools.py."""
    for s in lr.C_ABI_SYMBOLS:
        assert hasattr(oracle.lib, "lo_" + s)
    assert not any(hasattr(oracle.lib, "lo_" + s) for s in api.HOST_ONLY_SYMBOLS)


def test_create_fails_loudly_without_a_gpu_or_with_bad_settings():
    b = lr.bindings()
    h = ctypes.c_void_p()
    bad = lr.Settings(width=0, height=16).to_c()
    assert b.create(ctypes.byref(bad), ctypes.byref(h)) == -1
    import torch
    if not torch.cuda.is_available():
        ok = lr.Settings(width=16, height=16).to_c()
        rc = b.create(ctypes.byref(ok), ctypes.byref(h))
        assert rc == -3 and b"fallback" in b.last_error()      # LB_ERR_CUDA: there is no CPU path to fall back to
        with pytest.raises(lr.LumenError):
            lr.Renderer(width=16, height=16)
