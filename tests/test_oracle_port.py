"""The CPU restatement (oracle/ugport.c) against the golden dumps of the compiled reference.

Everything is required to be BIT-exact, reductions included: the restatement is sequential and
follows the reference's statement order, and both are compiled with -ffp-contract=off.
"""
import numpy as np
import pytest

from oracle.ugport import PortBackend
from replay import has, replay_krylov, replay_ops, replay_solve


def test_port_ops_bitexact(golden):
    if not has(golden, "ops"):
        pytest.skip("dump without per-call records")
    be = PortBackend(golden)
    n = replay_ops(be, golden, exact=True, exact_red=True)
    assert n > 10


def test_port_cycle_and_solve_bitexact(golden):
    if not has(golden, "solve"):
        pytest.skip("dump without solve records")
    be = PortBackend(golden)
    n = replay_solve(be, golden, exact=True, red_tol=0.0)
    assert n > 10


def test_port_krylov_bitexact(golden):
    """cg (ls.cc:989) and bcgs (ls.cc:1864) around the cycle: iterates, defects and histories bit for bit."""
    if not has(golden, "krylov"):
        pytest.skip("dump without Krylov records")
    be = PortBackend(golden)
    n = replay_krylov(be, golden, exact=True)
    assert n > 20


def test_port_galerkin_bitexact(golden):
    """AssembleGalerkinByMatrix (transgrid.cc:1575), cascaded from the top level down like the dump: every value of every Galerkin
    coarse-level matrix bit for bit; the product stays on the coarse pattern of these nested hierarchies."""
    d = golden.raw
    if "L0/galerkin/val" not in d:
        pytest.skip("dump without Galerkin records")
    be = PortBackend(golden)
    val = golden.levels[golden.top].val
    for l in range(golden.top, 0, -1):
        lc = golden.levels[l - 1]
        assert np.array_equal(d[f"L{l-1}/galerkin/rowptr"], lc.rowptr) and np.array_equal(d[f"L{l-1}/galerkin/col"], lc.col)
        val = be.galerkin(l, val)
        ref = d[f"L{l-1}/galerkin/val"]
        assert np.array_equal(val, ref), (l, int(np.count_nonzero(val != ref)), ref.size)
        assert np.count_nonzero(val) > 0


def fe_of(golden):
    d = golden.raw
    return dict(problem=0 if golden.bs == 1 else 1, dim=golden.dim, E=float(d["asm/E"][0]), nu=float(d["asm/nu"][0]), source=list(d["asm/source"]))


def test_port_assemble_bitexact(golden):
    """Element-loop assembly against what the reference's LocalAssemble (np/procs/assemble.cc:657) and
    AssembleDirichletBoundary (np/udm/disctools.cc:1837) leave with the element kernel of oracle/ug_driver.cc's class `fe`: every
    matrix value and the right-hand side of every level, bit for bit."""
    d = golden.raw
    if "L0/asm/val" not in d:
        pytest.skip("dump without assembly records")
    be = PortBackend(golden)
    for l in range(golden.top + 1):
        g = lambda k: d[f"L{l}/{k}"]
        val, b = be.assemble(l, fe_of(golden), g("elem_ptr"), g("elem_nodes"), g("asm/coef"), g("xyz"), g("asm/skip"), g("asm/sol"))
        assert np.array_equal(val, g("asm/val")), (l, int(np.count_nonzero(val != g("asm/val"))))
        assert np.array_equal(b, g("asm/rhs")), l
        assert np.count_nonzero(g("asm/skip")) > 0 and np.count_nonzero(b) > 0


def test_golden_invariants(golden):
    """Invariants the reference's own checkers assert (np/algebra/npcheck.cc:118-154)."""
    for l, lv in enumerate(golden.levels):
        assert lv.rowptr[0] == 0 and lv.rowptr[-1] == lv.col.size
        assert np.array_equal(lv.col[lv.rowptr[:-1]], np.arange(lv.n)), "diagonal first"
        new_defect = (lv.ctl & 1) != 0
        fine_dof = (lv.ctl & 2) != 0
        assert np.array_equal(new_defect, lv.vclass >= 2)
        assert np.array_equal(fine_dof, (lv.vclass >= 2) & (lv.vnclass <= 1))
        if l > 0:
            assert lv.p_rowptr[-1] == lv.p_col.size == lv.p_w.size
            # R is P restricted to fine rows with VCLASS >= NEWDEF_CLASS, transposed
            keep = np.repeat(lv.vclass >= 2, np.diff(lv.p_rowptr))
            assert lv.r_col.size == int(keep.sum())
            rows = np.repeat(np.arange(lv.n), np.diff(lv.p_rowptr))[keep]
            pt = sorted(zip(lv.p_col[keep].tolist(), rows.tolist(), lv.p_w[keep].tolist()))
            rr = np.repeat(np.arange(golden.levels[l - 1].n), np.diff(lv.r_rowptr))
            rt = sorted(zip(rr.tolist(), lv.r_col.tolist(), lv.r_w.tolist()))
            assert pt == rt
            # partition of unity of the P1/Q1 interpolation weights (geometric levels; the algebraic levels of an AMG transfer --
            # UG's levels < 1, the first -bottomlevel levels of such a dump -- carry Ruge-Stueben / Vanek weights)
            if l > (-int(golden.raw["bottomlevel"][0]) if "bottomlevel" in golden.raw else 0):
                s = np.add.reduceat(lv.p_w, lv.p_rowptr[:-1])
                assert np.allclose(s, 1.0, atol=1e-14)


def dirichlet_rows(val, rowptr, skip, bs):
    """What AssembleDirichletBoundary (np/udm/disctools.cc:1837) does to the matrix rows of vectors with skip bits: row j of every block
    of the vector's row zeroed, the diagonal entry (j, j) set to 1."""
    v = np.array(val, dtype=np.float64).reshape(-1, bs * bs)
    for r in np.nonzero(skip)[0]:
        for j in range(bs):
            if (int(skip[r]) >> j) & 1:
                v[rowptr[r]:rowptr[r + 1], j * bs:(j + 1) * bs] = 0.0
                v[rowptr[r], j * bs + j] = 1.0
    return v.reshape(-1)


def amg_levels(golden):
    return -int(golden.raw["bottomlevel"][0]) if "bottomlevel" in golden.raw else 0


def test_port_galerkin_pattern_growth_bitexact(golden):
    """The algebraic levels of the amg_* dumps were built by the reference's AMG transfer: every coarse matrix is AssembleGalerkinByMatrix
    (transgrid.cc:1575) on a level that held diagonal entries only, so all its connections were created by the product
    (CreateExtraConnection -> gm/algebra.cc:969).  The restatement must reproduce the PATTERN including the order of the rows' lists
    (diagonal, connections in reverse order of creation) and, cascaded from level 0 down like amgtransfer.cc:812-925, every value --
    the dump holds them after AssembleDirichletBoundary (amgtransfer.cc:1041), the cascade continues with the raw product."""
    namg = amg_levels(golden)
    if namg == 0:
        pytest.skip("dump without algebraic levels")
    be = PortBackend(golden)
    val = golden.levels[namg].val
    for k in range(namg, 0, -1):
        lc = golden.levels[k - 1]
        rp, col = be.galerkin_pattern(k)
        assert np.array_equal(rp, lc.rowptr) and np.array_equal(col, lc.col), k
        val = be.galerkin(k, val)
        assert np.array_equal(dirichlet_rows(val, lc.rowptr, lc.skip, lc.bs), lc.val), k
        # growing an EXISTING pattern: start from the diagonal and a symmetric subset of the connections -- the created ones must come
        # right after the diagonal, the old entries keep their order at the end of the row
        srp = [0]; scol = []
        for r in range(lc.n):
            row = lc.col[lc.rowptr[r]:lc.rowptr[r + 1]]
            scol += [int(row[0])] + [int(c) for c in row[1:] if (r + int(c)) % 3 == 0]
            srp.append(len(scol))
        rp2, col2 = be.galerkin_pattern(k, start=(np.array(srp, np.int32), np.array(scol, np.int32)))
        assert np.array_equal(rp2, lc.rowptr)
        for r in range(lc.n):
            nold = srp[r + 1] - srp[r] - 1
            got = col2[rp2[r]:rp2[r + 1]]
            assert got[0] == r and sorted(got) == sorted(lc.col[lc.rowptr[r]:lc.rowptr[r + 1]])
            assert list(got[len(got) - nold:]) == scol[srp[r] + 1:srp[r + 1]]


def test_product_galerkin_pattern_host_function(golden):
    """uggpu_galerkin_pattern (host half of uggpu_galerkin's pattern growth; no device involved) against the reference's patterns."""
    namg = amg_levels(golden)
    if namg == 0:
        pytest.skip("dump without algebraic levels")
    import ctypes as C
    from ug_b200 import capi
    L = capi.lib()
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in range(namg, 0, -1):
        lf, lc = golden.levels[k], golden.levels[k - 1]
        ar, ac, pr, pc = i32(lf.rowptr), i32(lf.col), i32(lf.p_rowptr), i32(lf.p_col)
        rp = np.zeros(lc.n + 1, np.int32)
        args = [C.c_int(lf.n), C.c_int(lc.n), p(ar), p(ac), p(pr), p(pc), None, None]
        assert L.uggpu_galerkin_pattern(*args, p(rp), None) == 0
        col = np.zeros(int(rp[-1]), np.int32)
        assert L.uggpu_galerkin_pattern(*args, p(rp), p(col)) == 0
        assert np.array_equal(rp, lc.rowptr) and np.array_equal(col, lc.col), k


def amg_config(golden):
    """(class, init string) of the AMG transfer numproc that built the dump's algebraic levels."""
    d = golden.raw
    if "amg/class" not in d:
        return None, ""
    return bytes(d["amg/class"]).decode(), bytes(d["amg/init"]).decode()


def is_rs(golden):
    cls, init = amg_config(golden)
    return cls == "selectionAMG" and "$C RugeStueben" in init and "$I RugeStueben" in init and "$strongRel 0.25" in init


def vanek_config(golden):
    """(theta, smooth) when the dump's algebraic levels were built by clusterAMG with VanekNeuss aggregation, else None."""
    import re
    cls, init = amg_config(golden)
    if cls != "clusterAMG" or "$C VanekNeuss" not in init:
        return None
    m = re.search(r"\$strongVanek ([0-9.eE+-]+)", init)
    smooth = 1 if "$I Vanek" in init else (0 if "$I PiecewiseConstant" in init else None)
    return (float(m.group(1)), smooth) if m and smooth is not None else None


def test_product_amg_vanek_host_function(golden):
    """uggpu_amg_vanek_host (MarkVanek + CoarsenVanek / GenerateClusters + IpVanek or IpPiecewiseConstant on the flat matrix) against the
    levels the reference's clusterAMG built: the same clusters in the same order and the same interpolation rows -- the cluster's own entry
    first, the smoothed ones behind it in the reference's list order, weights bit for bit."""
    cfg = vanek_config(golden)
    if cfg is None:
        pytest.skip("dump without Vanek aggregation levels")
    theta, smooth = cfg
    import ctypes as C
    from ug_b200 import capi
    L = capi.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    namg = amg_levels(golden)
    for k in range(namg, 0, -1):
        lf, lc = golden.levels[k], golden.levels[k - 1]
        n, nnz = lf.n, lf.col.size
        rp = np.ascontiguousarray(lf.rowptr, np.int32); col = np.ascontiguousarray(lf.col, np.int32)
        val = np.ascontiguousarray(lf.val, np.float64); skip = np.ascontiguousarray(lf.skip, np.uint32)
        cluster = np.zeros(n, np.int32); seed = np.zeros(n, np.int32)
        prp = np.zeros(n + 1, np.int32); pcol = np.zeros(nnz + n, np.int32); pw = np.zeros(nnz + n)
        nc = C.c_int(0)
        assert L.uggpu_amg_vanek_host(C.c_int(n), p(rp), p(col), p(val), p(skip), C.c_double(theta), C.c_int(smooth), p(cluster), p(seed),
                                      p(prp), p(pcol), p(pw), C.byref(nc)) == 0
        assert nc.value == lc.n, k
        z = int(prp[-1])
        assert np.array_equal(prp, lf.p_rowptr) and np.array_equal(pcol[:z], lf.p_col) and np.array_equal(pw[:z], lf.p_w), k
        has = np.diff(lf.p_rowptr) > 0
        assert np.array_equal(cluster[has], lf.p_col[lf.p_rowptr[:-1][has]]) and np.all(cluster[~has] == -1)
        assert np.all(lc.vclass == 3) and np.array_equal(lc.vnclass, lf.vclass[seed[:lc.n]]) and np.all(lc.ctl == 1) and np.all(lc.skip == 0)


def test_product_amg_rs_host_function(golden):
    """uggpu_amg_rs_host (MarkRelative + CoarsenRugeStueben + IpRugeStueben on the flat matrix; host half of uggpu_amg_coarsen_rs, no
    device involved) against the levels the reference's selectionAMG built: the same coarse points and the same interpolation rows --
    columns in the reference's list order, weights bit for bit -- on every algebraic level."""
    if not is_rs(golden):
        pytest.skip("dump without Ruge-Stueben levels")
    import ctypes as C
    from ug_b200 import capi
    L = capi.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    namg = amg_levels(golden)
    assert namg >= 2
    for k in range(namg, 0, -1):
        lf, lc = golden.levels[k], golden.levels[k - 1]
        n, nnz = lf.n, lf.col.size
        rp = np.ascontiguousarray(lf.rowptr, np.int32); col = np.ascontiguousarray(lf.col, np.int32)
        val = np.ascontiguousarray(lf.val, np.float64); skip = np.ascontiguousarray(lf.skip, np.uint32)
        coarse = np.zeros(n, np.uint8); prp = np.zeros(n + 1, np.int32); pcol = np.zeros(nnz + n, np.int32); pw = np.zeros(nnz + n)
        nc = C.c_int(0)
        assert L.uggpu_amg_rs_host(C.c_int(n), p(rp), p(col), p(val), p(skip), C.c_double(0.25), p(coarse), p(prp), p(pcol), p(pw), C.byref(nc)) == 0
        assert nc.value == lc.n == int(coarse.sum()), k
        z = int(prp[-1])
        assert np.array_equal(prp, lf.p_rowptr) and np.array_equal(pcol[:z], lf.p_col) and np.array_equal(pw[:z], lf.p_w), k
        # the coarse points are exactly the rows that interpolate from one vector with weight 1 and carry no skip bits' exception
        one = np.diff(lf.p_rowptr) == 1
        assert np.all(one[coarse == 1])
        # flags the new level must carry (GenerateNewGrid amgtools.cc:585-600)
        assert np.all(lc.vclass == 3) and np.array_equal(lc.vnclass, lf.vclass[coarse == 1]) and np.all(lc.ctl == 1)
        assert np.array_equal(lc.skip, lf.skip[coarse == 1])
