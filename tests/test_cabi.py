"""CPU tests: the C-ABI library loads and exports every symbol include/shadow_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "shadow_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(shadow_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from shadow_gnn_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/shadow_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert _lib.lib.shadow_version() >= 100


def test_no_cpu_fallback_without_device():
    """without a GPU the product refuses to run instead of silently computing on the host"""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from shadow_gnn_b200 import _lib
    import numpy as np
    h = ctypes.c_void_p()
    ip = np.array([0, 1, 2], np.uint32); ix = np.array([1, 0], np.uint32)
    rc = _lib.lib.shadow_sampler_create(ip.ctypes.data_as(ctypes.c_void_p), ix.ctypes.data_as(ctypes.c_void_p), 2, 2, b"", b"",
                                        4, 1, 0, 0, 1, ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in _lib.lib.shadow_last_error()
    import shadow_gnn_b200.ParallelSampler as PS
    with pytest.raises(_lib.ShadowError):
        PS.ParallelSampler(ip, ix, [], 4, 1, True, True, [], 1, "", "", "", 0)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under shadow_gnn_b200/ may reference it"""
    pkg = os.path.join(ROOT, "shadow_gnn_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
                assert "liboracle" not in src and "oracle/_ref" not in src, f"{fn} links the oracle"
