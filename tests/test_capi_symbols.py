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


def test_host_numprocs_bind_only_exported_symbols():
    """Every entry point the gpuls numproc family binds with dlsym (ug_b200/host/gpuls_np.cc UGGPU_FUNCS) is declared in include/uggpu.h and
    exported by libuggpu.so -- a missing one would make every numproc refuse to load on the GPU box -- and the CPU stand-in of the test suite
    (tests/standin) defines all of them too."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "ug_b200", "host", "gpuls_np.cc")).read()
    m = re.search(r"#define UGGPU_FUNCS\(X\)(.*?)\n\n", src, re.S)
    names = re.findall(r"X\((uggpu_[a-z0-9_]+)\)", m.group(1))
    assert len(names) >= 30
    declared = set(capi.declared_symbols())
    lib = capi.lib()
    standin = open(os.path.join(root, "tests", "standin", "uggpu_standin.cc")).read()
    for n in names:
        assert n in declared, n
        assert hasattr(lib, n), n
        assert re.search(r"\b" + n + r"\s*\(", standin), n
