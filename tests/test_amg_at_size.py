"""The host halves of the algebraic-level setup (uggpu_amg_rs_host, uggpu_amg_vanek_host, uggpu_galerkin_pattern: no device involved) at
sizes the committed fixtures do not reach: the unmodified reference (oracle/_ref, built where /root/reference exists) writes a dump with
its own AMG levels into a temporary directory, and every level is rebuilt from the one above it -- coarse points / clusters,
interpolation rows (list order, weights), Galerkin pattern (list order) -- and compared bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_oracle_port import amg_levels, dirichlet_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # Ruge-Stueben in 2D: 129^2 = 16 641 vectors on level 0, seven algebraic levels
    ("ugoracle2", ["--grid", "tri", "--refine", "7", "--collapse"], "selectionAMG", "$strongRel 0.25 $C RugeStueben $I RugeStueben $CM Galerkin $vectLimit 30 $hold", ("rs", 0.25, 0)),
    # Ruge-Stueben in 3D: 17^3 = 4 913 vectors; the Galerkin matrices densify (14 -> 43 entries per row on the second level)
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--collapse"], "selectionAMG", "$strongRel 0.25 $C RugeStueben $I RugeStueben $CM Galerkin $levelLimit -2 $hold", ("rs", 0.25, 0)),
    # aggregation, piecewise constant, 33^3 = 35 937 vectors
    ("ugoracle3", ["--grid", "tet", "--refine", "5", "--collapse"], "clusterAMG", "$strongVanek 0.08 $C VanekNeuss $I PiecewiseConstant $CM Galerkin $vectLimit 60 $hold", ("vanek", 0.08, 0)),
    # smoothed aggregation in 2D, 33^2 quadrilaterals (from 65^2 on the reference's own IpVanek follows a NULL interpolation matrix, amgtools.cc:3083)
    ("ugoracle2", ["--grid", "quad", "--refine", "5", "--collapse"], "clusterAMG", "$strongVanek 0.08 $C VanekNeuss $I Vanek $CM Galerkin $vectLimit 30 $hold", ("vanek", 0.08, 1)),
]
IDS = ["rs-tri-129^2", "rs-tet-17^3", "vanek-pc-tet-33^3", "vanek-quad-33^2"]


@pytest.mark.parametrize("exe,grid,cls,init,kind", CASES, ids=IDS)
def test_amg_host_setup_at_size(tmp_path, exe, grid, cls, init, kind):
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    from oracle.ugport import PortBackend
    from ug_b200 import capi
    from ug_b200.hierarchy import Hierarchy
    dump = str(tmp_path / "amg.ugh")
    out = subprocess.run([path] + grid + ["--amg", cls, init, "--lean", "--dump", dump], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-500:] + out.stderr[-500:]
    h = Hierarchy.from_ugh(dump)
    namg = amg_levels(h)
    assert namg >= 2
    L = capi.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    port = PortBackend(h)
    val = h.levels[namg].val
    for k in range(namg, 0, -1):
        lf, lc = h.levels[k], h.levels[k - 1]
        n, nnz = lf.n, lf.col.size
        rp = np.ascontiguousarray(lf.rowptr, np.int32); col = np.ascontiguousarray(lf.col, np.int32)
        v = np.ascontiguousarray(val, np.float64); skip = np.ascontiguousarray(lf.skip, np.uint32)
        prp = np.zeros(n + 1, np.int32); pcol = np.zeros(nnz + n, np.int32); pw = np.zeros(nnz + n)
        nc = C.c_int(0)
        if kind[0] == "rs":
            coarse = np.zeros(n, np.uint8)
            rc = L.uggpu_amg_rs_host(C.c_int(n), p(rp), p(col), p(v), p(skip), C.c_double(kind[1]), p(coarse), p(prp), p(pcol), p(pw), C.byref(nc))
        else:
            cluster = np.zeros(n, np.int32)
            rc = L.uggpu_amg_vanek_host(C.c_int(n), p(rp), p(col), p(v), p(skip), C.c_double(kind[1]), C.c_int(kind[2]), p(cluster), None, p(prp), p(pcol), p(pw), C.byref(nc))
        assert rc == 0 and nc.value == lc.n, (k, rc, nc.value, lc.n)
        z = int(prp[-1])
        assert np.array_equal(prp, lf.p_rowptr) and np.array_equal(pcol[:z], lf.p_col) and np.array_equal(pw[:z], lf.p_w), k
        # the Galerkin pattern the product creates on the fresh level, in the order of UG's lists
        pc = np.ascontiguousarray(pcol[:z])
        crp = np.zeros(lc.n + 1, np.int32)
        assert L.uggpu_galerkin_pattern(C.c_int(n), C.c_int(lc.n), p(rp), p(col), p(prp), p(pc), None, None, p(crp), None) == 0
        ccol = np.zeros(int(crp[-1]), np.int32)
        assert L.uggpu_galerkin_pattern(C.c_int(n), C.c_int(lc.n), p(rp), p(col), p(prp), p(pc), None, None, p(crp), p(ccol)) == 0
        assert np.array_equal(crp, lc.rowptr) and np.array_equal(ccol, lc.col), k
        # values: the port's product on that pattern, cascaded with the raw values (the dump holds them after AssembleDirichletBoundary)
        val = port.galerkin(k, val)
        assert np.array_equal(dirichlet_rows(val, lc.rowptr, lc.skip, lc.bs), lc.val), k
