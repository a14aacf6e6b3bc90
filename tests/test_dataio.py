"""savedata / loaddata against files the reference's SaveData wrote (np/udm/data_io.cc:650; `ug_driver --savedata`, the
files travel inside the dump as byte records).  The host-only halves of the C-ABI (uggpu_data_write / uggpu_data_read: format only) run
without a GPU; uggpu_savedata / uggpu_loaddata (body gathered / scattered on the device) are the GPU tests."""
import ctypes as C
import os

import numpy as np
import pytest

from ug_b200 import capi


class General(C.Structure):          # uggpu_data_general
    _fields_ = [("ident", C.c_char_p), ("mgfile", C.c_char_p), ("time", C.c_double), ("dt", C.c_double), ("ndt", C.c_double),
                ("nparfiles", C.c_int), ("me", C.c_int), ("magic_cookie", C.c_int)]


def _strs(items):
    return (C.c_char_p * len(items))(*[s.encode() for s in items])


def _general(d):
    return General(b"---", b"saved_without_mg", -1.0, -1.0, -1.0, 1, 0, int(d["savedata/magic_cookie"][0]))


def _body(golden):
    """what SaveData writes after the header: per node ID the components of sol, then of rhs"""
    d = golden.raw
    idl, idr, bs = d["savedata/id_level"], d["savedata/id_row"], golden.bs
    sol = [d[f"L{l}/savedata/sol"].reshape(-1, bs) for l in range(golden.top + 1)]
    rhs = [d[f"L{l}/savedata/rhs"].reshape(-1, bs) for l in range(golden.top + 1)]
    return np.concatenate([np.concatenate([sol[l][r], rhs[l][r]]) for l, r in zip(idl, idr)])


def _comp_names(bs):
    return "uvw"[:bs] if bs > 1 else "u"


def _has(golden):
    if "savedata/file_bin" not in golden.raw:
        pytest.skip("dump without savedata records")


@pytest.mark.parametrize("mode", ["bin", "asc"])
def test_data_write_bytes_equal_reference_file(golden, mode, tmp_path):
    _has(golden)
    d = golden.raw
    L = capi.lib()
    body = _body(golden)
    path = str(tmp_path / f"x.ug.data.{mode}")
    names = [s.decode() if isinstance(s, bytes) else s for s in ("sol", "rhs")]
    ncomp = (C.c_int * 2)(golden.bs, golden.bs)
    g = _general(d)
    cn = d["savedata/compnames"].tobytes().decode()
    rc = L.uggpu_data_write(path.encode(), mode.encode(), C.byref(g), 2, ncomp, _strs(names), _strs([cn, cn]), C.c_int64(len(d["savedata/id_level"])), capi._p(body))
    assert rc == 0, L.uggpu_last_error()
    assert open(path, "rb").read() == d[f"savedata/file_{mode}"].tobytes()


@pytest.mark.parametrize("mode", ["bin", "asc"])
def test_data_read_reference_file(golden, mode, tmp_path):
    _has(golden)
    d = golden.raw
    L = capi.lib()
    path = str(tmp_path / f"ref.ug.data.{mode}")
    open(path, "wb").write(d[f"savedata/file_{mode}"].tobytes())
    g = General(); nvd = C.c_int(0); ncomp = (C.c_int * 8)(); ndata = C.c_int64(0)
    assert L.uggpu_data_read(path.encode(), C.byref(g), C.byref(nvd), ncomp, 8, C.byref(ndata), None, C.c_int64(0)) == 0, L.uggpu_last_error()
    assert (nvd.value, ncomp[0], ncomp[1], g.magic_cookie, g.mgfile, g.ident) == (2, golden.bs, golden.bs, int(d["savedata/magic_cookie"][0]), b"saved_without_mg", b"---")
    body = _body(golden)
    assert ndata.value == body.size
    got = np.zeros(body.size)
    assert L.uggpu_data_read(path.encode(), C.byref(g), C.byref(nvd), ncomp, 8, C.byref(ndata), capi._p(got), C.c_int64(got.size)) == 0
    want = body if mode == "bin" else np.array([float("%g" % v) for v in body])      # ASCII files keep 6 significant digits (low/bio.cc:246)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["bin", "asc"])
def test_gpu_savedata_loaddata(golden, mode, tmp_path):
    """Device vectors -> file: bytes equal the reference's file.  Reference's file -> device vectors: every value where SaveData took it from."""
    _has(golden)
    d = golden.raw
    from backends import GpuBackend
    be = GpuBackend(golden)
    ctx = be.ctx
    for l in range(golden.top + 1):
        be.put(l, "sol", d[f"L{l}/savedata/sol"]); be.put(l, "rhs", d[f"L{l}/savedata/rhs"])
    idl = np.ascontiguousarray(d["savedata/id_level"]); idr = np.ascontiguousarray(d["savedata/id_row"])
    vec = (C.c_int * 2)(ctx.handle("sol"), ctx.handle("rhs"))
    g = _general(d)
    cn = d["savedata/compnames"].tobytes().decode()
    path = str(tmp_path / f"dev.ug.data.{mode}")
    ctx.call("uggpu_savedata", path.encode(), mode.encode(), C.byref(g), 2, vec, _strs(["sol", "rhs"]), _strs([cn, cn]), C.c_int64(idl.size), capi._p(idl), capi._p(idr))
    assert open(path, "rb").read() == d[f"savedata/file_{mode}"].tobytes()
    # load the REFERENCE's file into fresh vectors; the second descriptor of the file is skipped (as with a NULL entry in LoadData's list)
    ref = str(tmp_path / f"ref.ug.data.{mode}")
    open(ref, "wb").write(d[f"savedata/file_{mode}"].tobytes())
    for l, lv in enumerate(golden.levels):
        be.put(l, "a", np.full(lv.n * lv.bs, 9.0)); be.put(l, "b2", np.full(lv.n * lv.bs, 9.0))
    vec2 = (C.c_int * 2)(ctx.handle("a"), -1)
    ctx.call("uggpu_loaddata", ref.encode(), 2, vec2, C.c_int64(idl.size), capi._p(idl), capi._p(idr), None)
    rnd = (lambda v: v) if mode == "bin" else (lambda v: np.array([float("%g" % x) for x in v]))
    for l in range(golden.top + 1):
        assert np.array_equal(be.get(l, "a"), rnd(d[f"L{l}/savedata/sol"])), l
        assert np.all(be.get(l, "b2") == 9.0)
    be.close()
