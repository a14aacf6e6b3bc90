"""Properties of the algebraic-level setup that hold at any size (no reference dump involved): a hierarchy built from a synthetic 5-point
Laplacian by the product's host halves alone (uggpu_amg_rs_host / uggpu_amg_vanek_host + uggpu_galerkin_pattern; Galerkin values and the
cycle by the oracle's restatement) must (a) preserve constants where no Dirichlet row is involved (rows of P sum to 1), (b) give symmetric
Galerkin matrices with zero row sums away from the boundary, and (c) converge as a V(2,2) cycle at a rate that does not degrade with the size."""
import ctypes as C

import numpy as np
import pytest

from ug_b200 import capi
from ug_b200.hierarchy import Hierarchy, Level


def laplace2d(m):
    """5-point Laplacian on an m x m grid, diagonal first in every row, boundary rows = identity rows with skip bit (UG's Dirichlet treatment);
    interior rows keep their couplings to boundary nodes (columns of Dirichlet vectors), as an assembled UG matrix does."""
    n = m * m
    idx = lambda i, j: i * m + j
    rowptr = [0]; col = []; val = []
    skip = np.zeros(n, np.uint32)
    for i in range(m):
        for j in range(m):
            r = idx(i, j)
            if i in (0, m - 1) or j in (0, m - 1):
                skip[r] = 1
                col.append(r); val.append(1.0)
                for (a, b) in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1)):      # the connections exist in UG, their values are zeroed
                    if 0 <= a < m and 0 <= b < m:
                        col.append(idx(a, b)); val.append(0.0)
            else:
                col.append(r); val.append(4.0)
                for (a, b) in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1)):
                    col.append(idx(a, b)); val.append(-1.0)
            rowptr.append(len(col))
    return n, np.array(rowptr, np.int32), np.array(col, np.int32), np.array(val, np.float64), skip


def build(m, kind):
    L = capi.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    from oracle import ugport as up
    n, rp, col, val, skip = laplace2d(m)
    vclass = np.full(n, 3, np.uint8)
    levels = [Level(n=n, bs=1, rowptr=rp, col=col, val=val, vclass=vclass, vnclass=np.zeros(n, np.uint8), ctl=np.ones(n, np.uint8), skip=skip)]
    lib = None
    while levels[0].n > 40 and len(levels) < 12:
        lf = levels[0]
        n, nnz = lf.n, lf.col.size
        prp = np.zeros(n + 1, np.int32); pcol = np.zeros(nnz + n, np.int32); pw = np.zeros(nnz + n); nc = C.c_int(0)
        if kind == "rs":
            coarse = np.zeros(n, np.uint8)
            assert L.uggpu_amg_rs_host(C.c_int(n), p(lf.rowptr), p(lf.col), p(lf.val), p(lf.skip), C.c_double(0.25), p(coarse), p(prp), p(pcol), p(pw), C.byref(nc)) == 0
            cskip = lf.skip[coarse == 1].copy()
        else:
            cluster = np.zeros(n, np.int32)
            # piecewise constant interpolation: the smoothed one needs every free vector in a cluster, which the aggregation does not guarantee on
            # coarse Galerkin matrices (the reference follows a NULL pointer there, the product returns an error)
            assert L.uggpu_amg_vanek_host(C.c_int(n), p(lf.rowptr), p(lf.col), p(lf.val), p(lf.skip), C.c_double(0.08), C.c_int(0), p(cluster), None, p(prp), p(pcol), p(pw), C.byref(nc)) == 0
            cskip = np.zeros(nc.value, np.uint32)
        if nc.value == 0 or nc.value == n:
            break
        z = int(prp[-1]); pcol = np.ascontiguousarray(pcol[:z]); pw = np.ascontiguousarray(pw[:z])
        ncv = nc.value
        crp = np.zeros(ncv + 1, np.int32)
        assert L.uggpu_galerkin_pattern(C.c_int(n), C.c_int(ncv), p(lf.rowptr), p(lf.col), p(prp), p(pcol), None, None, p(crp), None) == 0
        ccol = np.zeros(int(crp[-1]), np.int32)
        assert L.uggpu_galerkin_pattern(C.c_int(n), C.c_int(ncv), p(lf.rowptr), p(lf.col), p(prp), p(pcol), None, None, p(crp), p(ccol)) == 0
        # R = P^T with the fine rows in list order
        order = np.argsort(pcol, kind="stable")
        rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(prp))
        rrp = np.zeros(ncv + 1, np.int32); np.add.at(rrp, pcol + 1, 1); rrp = np.cumsum(rrp).astype(np.int32)
        lf.p_rowptr, lf.p_col, lf.p_w = prp, pcol, pw
        lf.r_rowptr, lf.r_col, lf.r_w = rrp, np.ascontiguousarray(rows[order]), np.ascontiguousarray(pw[order])
        lc = Level(n=ncv, bs=1, rowptr=crp, col=ccol, val=np.zeros(ccol.size), vclass=np.full(ncv, 3, np.uint8), vnclass=np.full(ncv, 3, np.uint8),
                   ctl=np.ones(ncv, np.uint8), skip=cskip)
        if lib is None:
            lib = up.PortBackend(Hierarchy(dim=2, bs=1, fullrefinelevel=0, levels=[lc, lf], raw={"transfer_mode": np.array([1])})).L
        fine = up._Level(n, 1, ilu=None, rowptr=up._p(lf.rowptr), col=up._p(lf.col), val=up._p(lf.val), vclass=None, vnclass=None, ctl=None, skip=None,
                         p_rowptr=up._p(prp), p_col=up._p(pcol), p_w=up._p(pw), r_rowptr=None, r_col=None, r_w=None)
        coarseL = up._Level(ncv, 1, ilu=None, rowptr=up._p(crp), col=up._p(ccol), val=None, vclass=None, vnclass=None, ctl=None, skip=None,
                            p_rowptr=None, p_col=None, p_w=None, r_rowptr=None, r_col=None, r_w=None)
        assert lib.ugport_galerkin(C.byref(fine), C.byref(coarseL), up._dp(lf.val), up._dp(lc.val)) == 0
        levels.insert(0, lc)
    top = len(levels) - 1
    return Hierarchy(dim=2, bs=1, fullrefinelevel=top, levels=levels, raw={"transfer_mode": np.array([1])})


@pytest.mark.parametrize("kind", ["rs", "vanek"])
def test_amg_hierarchy_properties_and_convergence(kind):
    from oracle.ugport import PortBackend
    rates = {}
    for m in (33, 65):
        h = build(m, kind)
        assert h.top >= 2
        for l in range(1, h.top + 1):
            lv, lc = h.levels[l], h.levels[l - 1]
            # (a) constants are interpolated exactly at vectors all of whose strong neighbourhood is free of Dirichlet vectors
            s = np.add.reduceat(lv.p_w, lv.p_rowptr[:-1][np.diff(lv.p_rowptr) > 0])
            if kind == "rs" and l == h.top:        # (coarser levels: most rows feel the eliminated boundary, their row sums are not zero any more)
                assert np.count_nonzero(np.abs(s - 1.0) < 1e-12) > 0.5 * s.size
            # (b) the Galerkin matrix is symmetric
            import scipy.sparse as sp
            A = sp.csr_matrix((lc.val.copy(), lc.col.copy(), lc.rowptr.copy()), shape=(lc.n, lc.n))      # copies: scipy sorts the indices in place
            assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
        be = PortBackend(h)
        top = h.top
        rng = np.random.default_rng(7)
        b = rng.standard_normal(h.levels[top].n); b[h.levels[top].skip != 0] = 0.0
        for l in range(top + 1):
            be.put(l, "x", np.zeros(h.levels[l].n)); be.put(l, "b", b if l == top else np.zeros(h.levels[l].n))
        cfg = dict(nu1=2, nu2=2, gamma=1, baselevel=0, smoother="jac", smooth_damp=0.8, cycle_damp=1.0, base_maxit=10, base_reduction=1e-8, base_abslimit=1e-30)
        its, first, hist = be.solve(top, "x", "b", cfg, 8)
        assert its == 8, (kind, m, its, first, hist, [lv.n for lv in h.levels])
        rate = (hist[-1] / first[0]) ** (1.0 / 8)
        rates[m] = rate
        assert rate < (0.45 if kind == "rs" else 0.75), (kind, m, rate, hist)
    # (c) the rate does not degrade with the size (within 0.1)
    assert rates[65] < rates[33] + 0.1, rates
