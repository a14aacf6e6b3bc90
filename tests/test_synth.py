"""Synthetic hierarchies (ug_b200/csrc/synth.cu): (1) the 2D generator reproduces, entry by entry up to the row
permutation, the hierarchy the unmodified reference builds (golden c1); (2) on synthetic 3D hierarchies the CUDA
path is bit-identical to the oracle port, at sizes the port finishes in seconds; (3) size-independent properties
at a larger size (constant-preserving prolongation, R = P^T, symmetric A on free rows, monotone defect history)."""
import ctypes as C
import os

import numpy as np
import pytest

from ug_b200 import capi
from ug_b200.hierarchy import Hierarchy

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _synth(cells, dim, top, kind=capi.SYNTH_P1_SIMPLEX):
    ctx = capi.Context(0)
    ctx.call("uggpu_synth_hierarchy", kind, cells, cells, cells if dim == 3 else 0, top, ctx.handle("A"))
    return ctx


@pytest.fixture(params=["0", "1000000000"], ids=["tma", "thread-per-row"])
def tma_rows(request):
    return request.param


def _coords(cells, dim, level):
    nn = cells * 2 ** level + 1
    idx = np.arange(nn ** dim)
    xs = [(idx // nn ** d) % nn for d in range(dim)]
    return np.stack(xs, 1), nn


@pytest.mark.parametrize("name,dim,kind", [("c1_tri2d_r4", 2, capi.SYNTH_P1_SIMPLEX), ("c4_hex3d_bs3_r2", 3, capi.SYNTH_Q1_ELASTICITY)],
                         ids=["P1-triangles", "Q1-elasticity-3x3"])
def test_synth_equals_reference_hierarchy(name, dim, kind):
    """The generator reproduces what the unmodified reference builds + assembles (golden dump), up to the row permutation
    and the order of the entries inside a row: pattern, values (1e-12: the reference sums element contributions in
    its own order), flags, P, R, rhs."""
    gold = Hierarchy.from_ugh(os.path.join(GOLD, name + ".ugh"))
    ctx = _synth(1, dim, gold.top, kind)
    syn = ctx.download_hierarchy(gold.top)
    bb = gold.bs * gold.bs
    perm = []     # perm[l][gold row] = synthetic row
    for l, (g, s) in enumerate(zip(gold.levels, syn.levels)):
        xy, nn = _coords(1, dim, l)
        gij = np.rint(g.xyz.reshape(-1, dim) * (nn - 1)).astype(int)
        p = sum(gij[:, d] * nn ** d for d in range(dim))
        assert sorted(p.tolist()) == list(range(s.n))
        perm.append(p)
        assert g.n == s.n and g.nnz == s.nnz
        for k in ("vclass", "vnclass", "ctl", "skip"):
            assert np.array_equal(getattr(g, k), getattr(s, k)[p]), (l, k)
        assert s.bs == g.bs
        gv, sv = g.val.reshape(-1, bb), s.val.reshape(-1, bb)
        ge = {(p[r], p[g.col[e]]): gv[e] for r in range(g.n) for e in range(g.rowptr[r], g.rowptr[r + 1])}
        se = {(r, s.col[e]): sv[e] for r in range(s.n) for e in range(s.rowptr[r], s.rowptr[r + 1])}
        assert ge.keys() == se.keys()
        assert max(np.max(np.abs(ge[k] - se[k])) for k in ge) < 1e-12 * np.max(np.abs(g.val))
        assert np.array_equal(s.col[s.rowptr[:-1]], np.arange(s.n))      # diagonal first
        ctx.call("uggpu_synth_rhs", l, ctx.handle("b"))
        assert np.allclose(ctx.get(l, "b").reshape(-1, g.bs)[p], g.rhs.reshape(-1, g.bs), rtol=1e-12, atol=1e-18)
        if l > 0:
            pc = perm[l - 1]
            for pre in ("p", "r"):
                grp, gc, gw = (getattr(g, pre + k) for k in ("_rowptr", "_col", "_w"))
                srp, sc, sw = (getattr(s, pre + k) for k in ("_rowptr", "_col", "_w"))
                rowp, colp = (p, pc) if pre == "p" else (pc, p)
                gs = {(rowp[r], colp[gc[e]]): gw[e] for r in range(grp.size - 1) for e in range(grp[r], grp[r + 1])}
                ss = {(r, sc[e]): sw[e] for r in range(srp.size - 1) for e in range(srp[r], srp[r + 1])}
                assert gs == ss, (l, pre)
    ctx.close()


@pytest.mark.parametrize("cells,dim,top,kind", [(2, 3, 3, 0), (1, 3, 4, 0), (3, 2, 4, 0), (2, 3, 3, 1), (1, 3, 3, 2), (3, 3, 2, 2), (2, 3, 4, 3), (3, 2, 4, 3)],
                         ids=["P1-3d-17^3", "P1-3d-17^3-base1", "P1-2d-49^2", "Q1-17^3", "elast-9^3", "elast-13^3", "P1-varcoef-33^3", "P1-varcoef-2d-49^2"])
def test_synth_solve_bitexact_vs_port(cells, dim, top, kind, monkeypatch, tma_rows):
    monkeypatch.setenv("UGGPU_TMA", "1"); monkeypatch.setenv("UGGPU_TMA_MIN_ROWS", tma_rows)     # "0": every scalar level runs the bulk-copy staged smoothing kernel
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    ctx = _synth(cells, dim, top, kind)
    hier = ctx.download_hierarchy(top)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    rhs = ctx.get(top, "b")
    ctx.close()
    cfg = dict(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=0.6 if dim == 3 else 0.8)
    out = []
    for be in (GpuBackend(hier, fused=1), GpuBackend(hier, fused=0), PortBackend(hier)):
        for l, lv in enumerate(hier.levels):
            be.put(l, "x", np.zeros(lv.n * lv.bs)); be.put(l, "b", rhs if l == top else np.zeros(lv.n * lv.bs))
        be.ls_defect(0, top, "x", "b")
        its, first, hist = be.solve(top, "x", "b", cfg, 6)
        out.append((its, hist, [be.get(l, "x") for l in range(top + 1)], [be.get(l, "b") for l in range(top + 1)]))
        if hasattr(be, "close"):
            be.close()
    ref = out[-1]
    assert ref[0] == 6 and ref[1][-1] < 0.5 * ref[1][hier.bs - 1]
    # also with coupled FREE rows on the base level (cells > 2): the device LU follows the reference's list order, scalar and block
    exact = True
    bscale = np.max(np.abs(rhs))
    for its, hist, xs, bs in out[:-1]:
        assert its == ref[0]
        assert np.max(np.abs(hist - ref[1]) / np.maximum(ref[1], 1e-300)) < (1e-12 if exact else 1e-10)
        for l in range(top + 1):
            if exact:
                assert np.array_equal(xs[l], ref[2][l]), l
                assert np.array_equal(bs[l], ref[3][l]), l
            else:
                assert np.max(np.abs(xs[l] - ref[2][l])) <= 1e-12 * max(np.max(np.abs(ref[2][l])), 1e-300), l
                assert np.max(np.abs(bs[l] - ref[3][l])) <= 1e-12 * bscale, l


@pytest.mark.parametrize("cells,dim,top,kind,smoother,damp", [(2, 3, 3, 0, "gs", 0.9), (2, 3, 4, 0, "sgs", 0.8), (1, 3, 3, 2, "sor", 1.1), (2, 3, 3, 1, "sgs", 0.9),
                                                              (3, 2, 5, 0, "sor", 1.2)],
                         ids=["P1-17^3-gs", "P1-33^3-sgs", "elast-9^3-sor", "Q1-17^3-sgs", "P1-2d-97^2-sor"])
def test_synth_gs_family_bitexact_vs_port(cells, dim, top, kind, smoother, damp):
    """Gauss-Seidel family on lexicographically ordered synthetic hierarchies -- long dependency chains (a 33^3
    Kuhn grid has 97 levels per sweep): the triangular solves alone, then the cycle with the class as smoother, bit for bit
    against the sequential oracle port."""
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    ctx = _synth(cells, dim, top, kind)
    hier = ctx.download_hierarchy(top)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    rhs = ctx.get(top, "b")
    ctx.close()
    bs = hier.bs
    rng = np.random.default_rng(7)
    d0 = np.round(rng.standard_normal(hier.levels[top].n * bs) * 1024) / 1024
    omega = [0.25 + 0.5 * i for i in range(3)]
    gpu, port = GpuBackend(hier, fused=0), PortBackend(hier)
    for upper, om in ((False, None), (True, None), (False, omega), (True, omega)):
        res = []
        for be in (gpu, port):
            be.put(top, "d", d0); be.put(top, "v", np.full(d0.size, 3.0))
            assert be.l_gs(top, "v", "d", upper=upper, omega=om) == 0
            res.append(be.get(top, "v"))
        assert np.array_equal(res[0], res[1]), (upper, om)
    lo, up = C.c_int(), C.c_int()
    gpu.ctx.call("uggpu_gs_levels", top, gpu.A, C.byref(lo), C.byref(up))
    nn = cells * 2 ** top
    assert lo.value == up.value and lo.value >= nn          # at least one level per grid line
    cfg = dict(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=damp, smoother=smoother)
    out = []
    for be in (gpu, port):
        for l, lv in enumerate(hier.levels):
            be.put(l, "x", np.zeros(lv.n * lv.bs)); be.put(l, "b", rhs if l == top else np.zeros(lv.n * lv.bs))
        be.ls_defect(0, top, "x", "b")
        its, first, hist = be.solve(top, "x", "b", cfg, 4)
        out.append((its, hist, [be.get(l, "x") for l in range(top + 1)], [be.get(l, "b") for l in range(top + 1)]))
    gpu.close()
    (ig, hg, xg, bg), (ip, hp, xp, bp) = out
    assert ig == ip == 4 and hp[-1] < 0.2 * hp[bs - 1]
    assert np.max(np.abs(hg - hp) / np.maximum(hp, 1e-300)) < 1e-12
    for l in range(top + 1):
        assert np.array_equal(xg[l], xp[l]), l
        assert np.array_equal(bg[l], bp[l]), l


@pytest.mark.parametrize("cells,dim,top,kind,beta,damp", [(2, 3, 4, 0, 0.0, 1.0), (2, 3, 3, 0, 0.35, 0.9), (1, 3, 3, 2, 0.1, 0.9), (2, 3, 3, 1, 0.2, 1.0),
                                                          (3, 2, 5, 0, 0.5, 1.0)],
                         ids=["P1-33^3-ilu", "P1-17^3-ilu-beta", "elast-9^3-ilu-beta", "Q1-17^3-ilu-beta", "P1-2d-97^2-ilu-beta"])
def test_synth_ilu_bitexact_vs_port(cells, dim, top, kind, beta, damp):
    """ILU smoother (l_ilubthdecomp + l_luiter, class ilu) on lexicographic synthetic hierarchies: the decomposed
    matrix (every stored value), one l_luiter, then the cycle with the class as smoother -- bit for bit against the sequential
    oracle port (which is pinned against the reference's own dumps, tests/golden/ilu_*.ugh)."""
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    ctx = _synth(cells, dim, top, kind)
    hier = ctx.download_hierarchy(top)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    rhs = ctx.get(top, "b")
    ctx.close()
    bs = hier.bs
    rng = np.random.default_rng(11)
    d0 = np.round(rng.standard_normal(hier.levels[top].n * bs) * 1024) / 1024
    gpu, port = GpuBackend(hier, fused=0), PortBackend(hier)
    res = []
    for be in (gpu, port):
        assert be.ilu_decomp(top, beta) == 0
        be.put(top, "d", d0); be.put(top, "v", np.full(d0.size, 3.0))
        assert be.l_luiter(top, "v", "d") == 0
        res.append((be.ilu_values(top), be.get(top, "v")))
    assert np.array_equal(res[0][0], res[1][0]), "decomposition differs"
    assert np.array_equal(res[0][1], res[1][1]), "l_luiter differs"
    assert not np.array_equal(res[1][0], hier.levels[top].val)
    cfg = dict(nu1=1, nu2=1, gamma=1, baselevel=0, smooth_damp=damp, smoother="ilu", ilu_beta=beta)
    out = []
    for be in (gpu, port):
        for l, lv in enumerate(hier.levels):
            be.put(l, "x", np.zeros(lv.n * lv.bs)); be.put(l, "b", rhs if l == top else np.zeros(lv.n * lv.bs))
        be.ls_defect(0, top, "x", "b")
        its, first, hist = be.solve(top, "x", "b", cfg, 4)
        out.append((its, hist, [be.get(l, "x") for l in range(top + 1)], [be.get(l, "b") for l in range(top + 1)]))
    gpu.close()
    (ig, hg, xg, bg), (ip, hp, xp, bp) = out
    assert ig == ip == 4 and hp[-1] < 0.2 * hp[bs - 1]
    assert np.max(np.abs(hg - hp) / np.maximum(hp, 1e-300)) < 1e-12
    for l in range(top + 1):
        assert np.array_equal(xg[l], xp[l]), l
        assert np.array_equal(bg[l], bp[l]), l



def test_synth_properties_large():
    """65^3 = 274 625 unknowns: properties that do not need the CPU checker."""
    cells, top = 2, 5
    ctx = _synth(cells, 3, top)
    n = ctx.level_n(top)
    assert n == 65 ** 3
    A = ctx.handle("A")
    one = capi._vs([1.0])
    # prolongation preserves constants on free rows, is zero on Dirichlet rows
    ctx.put(top - 1, "c", np.ones(ctx.level_n(top - 1)))
    ctx.alloc(top, "t")
    ctx.call("uggpu_interpolate_correction", top, ctx.handle("t"), ctx.handle("c"), one)
    t = ctx.get(top, "t")
    skip = np.zeros(n, np.uint32)
    ctx.call("uggpu_level_get_flags", top, None, None, None, skip.ctypes.data_as(C.c_void_p))
    assert np.array_equal(t, (skip == 0).astype(float))
    # <R u, v> = <u, P v> for v supported on free coarse rows (adjointness of restriction and prolongation)
    rng = np.random.default_rng(7)
    u = rng.integers(-8, 8, n).astype(float) * (skip == 0)
    skc = np.zeros(ctx.level_n(top - 1), np.uint32)
    ctx.call("uggpu_level_get_flags", top - 1, None, None, None, skc.ctypes.data_as(C.c_void_p))
    v = rng.integers(-8, 8, skc.size).astype(float) * (skc == 0)
    ctx.put(top, "u", u); ctx.put(top - 1, "v", v); ctx.alloc(top - 1, "u"); ctx.alloc(top, "v")
    ctx.call("uggpu_restrict", top, ctx.handle("u"), ctx.handle("u"), one)
    ctx.call("uggpu_interpolate_correction", top, ctx.handle("v"), ctx.handle("v"), one)
    assert np.dot(ctx.get(top - 1, "u"), v) == np.dot(u, ctx.get(top, "v"))     # dyadic weights, small integers: exact
    # A is symmetric on the free rows: <A u, w> = <u, A w>
    w = rng.integers(-8, 8, n).astype(float) * (skip == 0)
    ctx.put(top, "w", w); ctx.alloc(top, "Au"); ctx.alloc(top, "Aw")
    ctx.call("uggpu_dmatmul", top, top, 0, ctx.handle("Au"), A, ctx.handle("u"))
    ctx.call("uggpu_dmatmul", top, top, 0, ctx.handle("Aw"), A, ctx.handle("w"))
    lhs, rhs = np.dot(ctx.get(top, "Au") * (skip == 0), w), np.dot(u, ctx.get(top, "Aw") * (skip == 0))
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)
    # V-cycles converge with a level-independent rate (h-independence of multigrid)
    for name in ("x", "b", "c"):
        for l in range(top + 1):
            ctx.alloc(l, name)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    cfg = ctx.lmgc_cfg(smooth_damp=0.6, fused=1)
    ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
    res = capi.LResult()
    ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
    hist = np.zeros(8)
    ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), 8,
             capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), hist.ctypes.data_as(C.POINTER(C.c_double)))
    rates = hist[1:] / hist[:-1]
    assert np.all(rates < 0.75) and res.number_of_linear_iterations == 8
    # the defect the solver reports is the defect of the iterate it returns: b0 - A x == b (up to rounding)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("w"))
    ctx.call("uggpu_dmatmul_minus", top, top, 0, ctx.handle("w"), A, ctx.handle("x"))
    assert np.max(np.abs(ctx.get(top, "w") - ctx.get(top, "b"))) < 1e-12 * np.max(np.abs(hist[0]))
    ctx.close()


def test_col_compression_lossless(monkeypatch):
    """Slices of rows with identical column distances store one word per slice column (sell.cu sell_compress_cols):
    the decoded pattern and every result are identical to the explicit storage, only fewer column words are read."""
    def run(compress):
        if compress:
            monkeypatch.delenv("UGGPU_NO_COL_COMPRESSION", raising=False)
        else:
            monkeypatch.setenv("UGGPU_NO_COL_COMPRESSION", "1")
        ctx = _synth(4, 3, 4)                      # 65^3: lines of 65 rows hold slices of 32 interior rows
        top, A = 4, ctx.handle("A")
        hier = ctx.download_hierarchy(top)
        words = int(ctx.L.uggpu_mat_col_words(ctx.h, top, A))
        nnz = int(ctx.L.uggpu_mat_nnz(ctx.h, top, A))
        for name in ("x", "b", "c"):
            for l in range(top + 1):
                ctx.alloc(l, name)
        ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
        cfg = ctx.lmgc_cfg(smooth_damp=0.6, fused=1)
        ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
        res = capi.LResult()
        ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
        ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), 3,
                 capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), None)
        x, b = ctx.get(top, "x"), ctx.get(top, "b")
        ctx.close()
        return hier, words, nnz, x, b

    h1, w1, nnz, x1, b1 = run(True)
    h0, w0, _, x0, b0 = run(False)
    assert w0 == nnz and w1 < 0.8 * nnz, (w0, w1, nnz)
    for a, b in zip(h1.levels, h0.levels):
        assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val)
    assert np.array_equal(x1, x0) and np.array_equal(b1, b0)


@pytest.mark.parametrize("kind,cells,top", [(capi.SYNTH_P1_SIMPLEX, 4, 4), (capi.SYNTH_Q1_ELASTICITY, 4, 4)], ids=["P1-65^3", "elast-65^3"])
def test_shared_value_tables_lossless(monkeypatch, kind, cells, top):
    """Uniform slices whose rows also hold bit-identical values read one shared value table (sell.cu sell_share_values): the
    matrix handed back by uggpu_mat_get and every result of a solve are identical to the explicit storage -- only the values
    fetched per sweep shrink.  After uggpu_mat_set_values with perturbed values the tables follow (or vanish)."""
    def run(share):
        if share:
            monkeypatch.delenv("UGGPU_NO_SHARED_VALUES", raising=False)
        else:
            monkeypatch.setenv("UGGPU_NO_SHARED_VALUES", "1")
        ctx = _synth(cells, 3, top, kind)
        A = ctx.handle("A")
        hier = ctx.download_hierarchy(top)
        ve = int(ctx.L.uggpu_mat_val_entries(ctx.h, top, A))
        nnz = int(ctx.L.uggpu_mat_nnz(ctx.h, top, A))
        for name in ("x", "b", "c", "y"):
            for l in range(top + 1):
                ctx.alloc(l, name)
        ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
        cfg = ctx.lmgc_cfg(smooth_damp=0.6, fused=1)
        ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
        res = capi.LResult()
        ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
        ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), 3,
                 capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), None)
        x, b = ctx.get(top, "x"), ctx.get(top, "b")
        # new values on the same pattern: one row in the middle changes -> its slice must fall back to explicit values
        lv = hier.levels[top]
        val = lv.val.copy()
        mid = lv.n // 2
        bb = lv.bs * lv.bs
        val[lv.rowptr[mid] * bb:lv.rowptr[mid + 1] * bb] *= 1.5
        ctx.call("uggpu_mat_set_values", top, A, capi._p(val))
        ve2 = int(ctx.L.uggpu_mat_val_entries(ctx.h, top, A))
        ctx.call("uggpu_dmatmul", top, top, 0, ctx.handle("y"), A, ctx.handle("x"))
        y = ctx.get(top, "y")
        back = ctx.download_hierarchy(top).levels[top].val
        ctx.close()
        return hier, ve, ve2, nnz, x, b, y, back, val

    h1, ve1, ve1b, nnz, x1, b1, y1, back1, val = run(True)
    h0, ve0, ve0b, _, x0, b0, y0, back0, _ = run(False)
    assert ve0 == nnz == ve0b and ve1 < 0.6 * nnz, (ve0, ve1, nnz)
    assert ve1 < ve1b < 0.6 * nnz, (ve1, ve1b)                     # the perturbed row's slice reads its own values again
    for a, b in zip(h1.levels, h0.levels):
        assert np.array_equal(a.rowptr, b.rowptr) and np.array_equal(a.col, b.col) and np.array_equal(a.val, b.val)
    assert np.array_equal(x1, x0) and np.array_equal(b1, b0)
    assert np.array_equal(back1, val) and np.array_equal(back0, val)
    assert np.array_equal(y1, y0)


@pytest.mark.parametrize("kind,top,minfrac", [(capi.SYNTH_P1_SIMPLEX, 5, None), (capi.SYNTH_Q1_POISSON, 5, None), (capi.SYNTH_Q1_ELASTICITY, 4, "0.3")],
                         ids=["P1-129^3-15pt", "Q1-129^3-27pt", "elast-65^3-27x3x3"])
def test_stencil_smoothing_kernel_bitexact_vs_port(monkeypatch, kind, top, minfrac):
    """129^3: most slices of the finest level carry one stencil, so the fused schedule runs the stencil variant of the smoothing
    kernel (spmv.cu k_smooth_sten / k_smooth_sten3 for 3x3 blocks: distance and value tables as kernel parameters).  Its solve must equal, bit for bit, the
    generic kernel's (UGGPU_NO_STENCIL=1), the one-kernel-per-call schedule's and the sequential oracle port's."""
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    if minfrac:       # 3x3 blocks at a size the port finishes in seconds (65^3 nodes: half of the slices are interior): a smaller share counts as dominant
        monkeypatch.setenv("UGGPU_STENCIL_MIN_FRAC", minfrac)
    ctx = _synth(4, 3, top, kind)
    hier = ctx.download_hierarchy(top)
    A = ctx.handle("A")
    sten = [int(ctx.L.uggpu_mat_stencil_slices(ctx.h, l, A)) for l in range(top + 1)]
    nsl = (hier.levels[top].n + 31) // 32
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    rhs = ctx.get(top, "b")
    ctx.close()
    assert sten[top] > (float(minfrac) if minfrac else 0.6) * nsl, sten   # the kernel under test is the one that runs on the finest level
    assert sten[0] == 0 and sten[1] == 0, sten                        # ... and the small levels keep the generic kernel
    cfg = dict(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=0.6)
    out = []
    for name in ("stencil", "slice-stencil", "generic", "no-row-classes", "per-call", "port"):
        # stencil: stencil rows + exception rows as two kernels (stx.cu); slice-stencil: the one-kernel form that decides per slice (UGGPU_NO_STX=1)
        if name == "generic":
            monkeypatch.setenv("UGGPU_NO_STENCIL", "1")
        else:
            monkeypatch.delenv("UGGPU_NO_STENCIL", raising=False)
        if name == "slice-stencil":
            monkeypatch.setenv("UGGPU_NO_STX", "1")
        else:
            monkeypatch.delenv("UGGPU_NO_STX", raising=False)
        if name == "no-row-classes":       # transfer.cu's general kernels instead of the row-class form (trc.cu)
            monkeypatch.setenv("UGGPU_NO_TRC", "1")
        else:
            monkeypatch.delenv("UGGPU_NO_TRC", raising=False)
        be = PortBackend(hier) if name == "port" else GpuBackend(hier, fused=0 if name == "per-call" else 1)
        for l, lv in enumerate(hier.levels):
            be.put(l, "x", np.zeros(lv.n * lv.bs)); be.put(l, "b", rhs if l == top else np.zeros(lv.n * lv.bs))
        be.ls_defect(0, top, "x", "b")
        its, first, hist = be.solve(top, "x", "b", cfg, 3)
        out.append((its, hist, [be.get(l, "x") for l in range(top + 1)], [be.get(l, "b") for l in range(top + 1)]))
        if hasattr(be, "close"):
            be.close()
    ref = out[-1]
    assert ref[0] == 3 and ref[1][-1] < 0.8 * ref[1][hier.bs - 1]        # the defect falls (same component, cycle 1 -> cycle 3)
    for its, hist, xs, bs in out[:-1]:
        assert its == 3
        assert np.max(np.abs(hist - ref[1]) / ref[1]) < 1e-12
        for l in range(top + 1):
            assert np.array_equal(xs[l], ref[2][l]), l
            assert np.array_equal(bs[l], ref[3][l]), l


def test_async_upload_of_the_iterate():
    """uggpu_vec_upload_async: the iterate travels on the copy stream while the cycle runs; results are those of the
    synchronous upload, also when the vector is reused at once (upload ordered behind the kernels already enqueued)."""
    ctx = _synth(2, 3, 4)
    top, A = 4, ctx.handle("A")
    for name in ("x", "b", "c"):
        for l in range(top + 1):
            ctx.alloc(l, name)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    b0 = ctx.get(top, "b")
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(b0.size)
    cfg = ctx.lmgc_cfg(smooth_damp=0.6, fused=1)
    ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
    out = []
    for mode in ("sync", "async", "async"):
        ctx.put(top, "b", b0)
        if mode == "sync":
            ctx.put(top, "x", x0)
        else:
            ctx.call("uggpu_vec_upload_async", top, ctx.handle("x"), x0.ctypes.data_as(C.c_void_p))
        res = capi.LResult()
        ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
        ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), 2,
                 capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), None)
        out.append((ctx.get(top, "x"), ctx.get(top, "b")))
    for x, b in out[1:]:
        assert np.array_equal(x, out[0][0]) and np.array_equal(b, out[0][1])
    # a pending upload is honoured by plain vector operations and downloads too
    ctx.call("uggpu_vec_upload_async", top, ctx.handle("x"), x0.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ctx.get(top, "x"), x0)
    ctx.close()
