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
