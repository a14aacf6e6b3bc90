"""The C-ABI library loads here (no GPU) and exports every function include/uggpu.h declares; without a device the
context constructor fails loudly instead of falling back to anything."""
import ctypes as C
import os

import pytest

from ug_b200 import capi


def _built():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.lib()


def test_exports_every_declared_symbol():
    L = _built()
    names = capi.declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _built()
    h = C.c_void_p()
    rc = L.uggpu_ctx_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in L.uggpu_last_error()
