"""CUDA path (through the C-ABI) against the golden dumps of the compiled reference.

Bar: every vector the reference produces is reproduced BIT-EXACTLY (both schedules: one kernel per reference
call, and the fused kernels); reductions (ddot/dnrm2 families, defect history) within 1e-13 / 1e-12 relative,
because a parallel sum cannot follow the reference's single running sum.
"""
import numpy as np
import pytest

from replay import has, replay_krylov, replay_ops, replay_solve

pytestmark = pytest.mark.gpu


@pytest.fixture
def gpu_backend(golden, request, monkeypatch):
    from backends import GpuBackend
    fused = getattr(request, "param", 1)
    if fused == 2:      # fused schedule with the bulk-copy (TMA) staged smoothing kernel forced onto every scalar level
        monkeypatch.setenv("UGGPU_TMA", "1"); monkeypatch.setenv("UGGPU_TMA_MIN_ROWS", "0")
        fused = 1
    if fused == 3:      # fused schedule; the row-class transfer (trc.cu) and the stencil-rows / exception-rows kernels (stx.cu) tried on every level, however small
        monkeypatch.setenv("UGGPU_TRC_MIN_ROWS", "1"); monkeypatch.setenv("UGGPU_STX_MIN_ROWS", "1"); monkeypatch.setenv("UGGPU_STENCIL_MIN_FRAC", "0.05")
        fused = 1
    be = GpuBackend(golden, fused=fused)
    yield be
    be.close()


def test_pattern_roundtrip_bitexact(golden):
    """uggpu_mat_set -> device SELL-32 -> uggpu_mat_get returns the canonical CSR/BSR and stencils unchanged."""
    from backends import GpuBackend
    be = GpuBackend(golden)
    back = be.ctx.download_hierarchy(golden.top)
    for l, (a, b) in enumerate(zip(golden.levels, back.levels)):
        for k in ("rowptr", "col", "val", "vclass", "vnclass", "ctl", "skip"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (l, k)
        if l > 0:
            for k in ("p_rowptr", "p_col", "p_w", "r_rowptr", "r_col", "r_w"):
                assert np.array_equal(getattr(a, k), getattr(b, k)), (l, k)
    be.close()


@pytest.mark.parametrize("gpu_backend", [1, 3], indirect=True, ids=["default", "row-forms-everywhere"])
def test_gpu_ops_bitexact(gpu_backend, golden):
    if not has(golden, "ops"):
        pytest.skip("dump without per-call records")
    n = replay_ops(gpu_backend, golden, exact=True, red_tol=1e-13)
    assert n > 10


@pytest.mark.parametrize("gpu_backend", [0, 1, 2, 3], indirect=True, ids=["per-call", "fused", "fused-tma", "fused-row-forms-everywhere"])
def test_gpu_cycle_and_solve(gpu_backend, golden):
    if not has(golden, "solve"):
        pytest.skip("dump without solve records")
    # Base levels with FREE rows (lu_* fixtures): the device follows the order of UG's matrix lists with fill-in, scalar and
    # block elimination alike (cycle.cu lu_lists) -- bit for bit.
    # transfer $L (lopt_* fixtures): the two scalars of MinimizeLevel are parallel sums on the device, so everything downstream of them
    # agrees to rounding (1e-11 of the largest entry), not bit for bit
    lopt = "level_opt" in golden.raw and int(golden.raw["level_opt"][0]) != 0
    n = replay_solve(gpu_backend, golden, exact=not lopt, vec_tol=1e-11, red_tol=1e-11 if lopt else 1e-12)
    assert n > 10


@pytest.mark.parametrize("gpu_backend", [0, 1], indirect=True, ids=["per-call", "fused"])
def test_gpu_krylov(gpu_backend, golden):
    """cg / bcgs of the reference around the cycle (ls.cc:989, :1864), device-resident.  The step lengths come from parallel
    sums, so iterates agree to rounding amplified by the iteration (1e-10 of the largest entry), histories to 1e-9 while they
    are above 1e-10 of the first defect; iteration counts exactly."""
    if not has(golden, "krylov"):
        pytest.skip("dump without Krylov records")
    n = replay_krylov(gpu_backend, golden, exact=False, vec_tol=1e-10, red_tol=1e-9)
    assert n > 20


def test_gpu_galerkin_bitexact(golden):
    """uggpu_galerkin against AssembleGalerkinByMatrix of the reference (transgrid.cc:1575), cascaded from the top level down like
    the dump: every value of every Galerkin coarse-level matrix bit for bit; afterwards a product with the new coarse matrix
    equals the port's (the derived storage forms -- diagonal array, shared value tables -- follow the new values)."""
    d = golden.raw
    if "L0/galerkin/val" not in d:
        pytest.skip("dump without Galerkin records")
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    be = GpuBackend(golden)
    for l in range(golden.top, 0, -1):
        val = be.galerkin(l)
        ref = d[f"L{l-1}/galerkin/val"]
        assert np.array_equal(val, ref), (l, int(np.count_nonzero(val != ref)), ref.size)
    # the coarse matrices now in use: x = A y on level top-1 against the port working on the dumped Galerkin values
    l = golden.top - 1
    lv = golden.levels[l]
    y = np.round(np.random.default_rng(3).standard_normal(lv.n * lv.bs) * 1024) / 1024
    be.put(l, "y", y); be.put(l, "x", np.zeros_like(y))
    be.dmatmul(l, l, 0, 0, "x", "y")
    got = be.get(l, "x")
    be.close()
    import copy
    h2 = copy.copy(golden); h2.levels = list(golden.levels); h2.levels[l] = copy.copy(lv); h2.levels[l].val = d[f"L{l}/galerkin/val"]
    port = PortBackend(h2)
    port.put(l, "y", y); port.put(l, "x", np.zeros_like(y))
    port.dmatmul(l, l, 0, 0, "x", "y")
    assert np.array_equal(got, port.get(l, "x"))


def test_gpu_assemble_bitexact(golden):
    """uggpu_assemble against the reference's LocalAssemble + AssembleDirichletBoundary (np/procs/assemble.cc:657,
    np/udm/disctools.cc:1837; dumps written by `ug_driver --assemble`): matrix values, right-hand side and VECSKIP of every level bit
    for bit -- into a matrix created from the pattern alone (uggpu_mat_set_pattern) -- and a product with the assembled matrix equals
    the port's on the dumped values (diagonal array and shared tables follow the new values)."""
    d = golden.raw
    if "L0/asm/val" not in d:
        pytest.skip("dump without assembly records")
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    from test_oracle_port import fe_of
    from ug_b200 import capi
    be = GpuBackend(golden)
    ctx = be.ctx
    for l, lv in enumerate(golden.levels):
        g = lambda k: d[f"L{l}/{k}"]
        ctx.call("uggpu_mat_set_pattern", l, ctx.handle("K"), capi._p(np.ascontiguousarray(lv.rowptr)), capi._p(np.ascontiguousarray(lv.col)))
        be.put(l, "x", g("asm/sol")); be.put(l, "b", np.full(lv.n * lv.bs, 7.0))
        ctx.assemble(l, "x", "b", "K", fe_of(golden), g("elem_ptr"), g("elem_nodes"), g("asm/coef"), g("xyz"), g("asm/skip"))
        val = ctx.mat_values(l, "K", lv.col.size)
        ref = g("asm/val")
        assert np.array_equal(val, ref), (l, int(np.count_nonzero(val != ref)), ref.size)
        assert np.array_equal(be.get(l, "b"), g("asm/rhs")), l
        assert np.array_equal(be.get(l, "x"), g("asm/sol")), l
        skip = np.zeros(lv.n, np.uint32)
        ctx.call("uggpu_level_get_flags", l, None, None, None, capi._p(skip))
        assert np.array_equal(skip, g("asm/skip"))
    l = golden.top
    lv = golden.levels[l]
    y = np.round(np.random.default_rng(5).standard_normal(lv.n * lv.bs) * 1024) / 1024
    be.put(l, "y", y); be.put(l, "z", np.zeros_like(y))
    ctx.call("uggpu_dmatmul", l, l, 0, ctx.handle("z"), ctx.handle("K"), ctx.handle("y"))
    got = be.get(l, "z")
    be.close()
    import copy
    h2 = copy.copy(golden); h2.levels = list(golden.levels); h2.levels[l] = copy.copy(lv); h2.levels[l].val = d[f"L{l}/asm/val"]
    port = PortBackend(h2)
    port.put(l, "y", y); port.put(l, "z", np.zeros_like(y))
    port.dmatmul(l, l, 0, 0, "z", "y")
    assert np.array_equal(got, port.get(l, "z"))


def test_gpu_krylov_fused_chains_bitexact(golden, monkeypatch):
    """cg / bcgs with the fused BLAS-1 chains (blas1.cu chain_loop: CGUpdate's calls in three passes, the bcgs updates in one pass each,
    the residuum's partial sums inside the pass that produces the defect) against the same solvers issuing one kernel per reference call
    (UGGPU_NO_KRYLOV_FUSION=1): iterates, defects, every work vector's effect and the defect history BIT for bit -- the chains perform
    the same operations per entry in the same order and the reductions keep launch geometry and accumulation order."""
    if not has(golden, "krylov"):
        pytest.skip("dump without Krylov records")
    from backends import GpuBackend
    from replay import cycle_cfg
    cfg = cycle_cfg(golden)
    top = golden.top
    out = {}
    for fused in (1, 0):
        if fused:
            monkeypatch.delenv("UGGPU_NO_KRYLOV_FUSION", raising=False)
        else:
            monkeypatch.setenv("UGGPU_NO_KRYLOV_FUSION", "1")
        be = GpuBackend(golden, fused=1)
        res = []
        for name in ("cg", "bcgs"):
            for l, lv in enumerate(golden.levels):
                be.put(l, "x", np.zeros(lv.n * lv.bs)); be.put(l, "b", lv.rhs)
            be.ls_defect(0, top, "x", "b")
            its, first, hist = (be.cg_solve if name == "cg" else be.bcgs_solve)(top, "x", "b", cfg, 5)
            res.append((its, first, hist, [be.get(l, "x") for l in range(top + 1)], [be.get(l, "b") for l in range(top + 1)]))
        out[fused] = res
        be.close()
    for a, b in zip(out[1], out[0]):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        for va, vb in zip(a[3] + a[4], b[3] + b[4]):
            assert np.array_equal(va, vb)


def test_gpu_galerkin_pattern_growth(golden):
    """uggpu_galerkin on levels whose coarse matrix does not exist yet (the algebraic levels of the amg_* dumps, built by the reference's
    AMG transfer with AssembleGalerkinByMatrix + CreateExtraConnection): the coarse PATTERN -- order of the rows' lists included -- and,
    cascaded from level 0 down, every VALUE bit for bit (the dump holds the values after AssembleDirichletBoundary; the raw products are
    compared with the port's, which the CPU suite pins against the dump).  Then a product that has to GROW an existing pattern."""
    from test_oracle_port import amg_levels, dirichlet_rows
    namg = amg_levels(golden)
    if namg == 0:
        pytest.skip("dump without algebraic levels")
    from backends import GpuBackend
    from oracle.ugport import PortBackend
    from ug_b200 import capi
    be = GpuBackend(golden)
    port = PortBackend(golden)
    ctx = be.ctx
    pval = golden.levels[namg].val
    for k in range(namg, 0, -1):
        lc = golden.levels[k - 1]
        ctx.call("uggpu_mat_free", k - 1, be.A)
        val = be.galerkin(k)                       # asserts the pattern against the dump's
        pval = port.galerkin(k, pval)
        assert np.array_equal(val, pval), (k, int(np.count_nonzero(val != pval)), pval.size)
        assert np.array_equal(dirichlet_rows(val, lc.rowptr, lc.skip, lc.bs), lc.val), k
    # growth of an existing pattern: level namg-1 restarts from the diagonal and a symmetric subset of its connections
    k = namg
    lc = golden.levels[k - 1]
    srp = [0]; scol = []
    for r in range(lc.n):
        row = lc.col[lc.rowptr[r]:lc.rowptr[r + 1]]
        scol += [int(row[0])] + [int(c) for c in row[1:] if (r + int(c)) % 3 == 0]
        srp.append(len(scol))
    srp = np.array(srp, np.int32); scol = np.array(scol, np.int32)
    ctx.call("uggpu_mat_set_pattern", k - 1, be.A, capi._p(srp), capi._p(scol))
    ctx.call("uggpu_galerkin", k, be.A)
    rp2, col2 = port.galerkin_pattern(k, start=(srp, scol))
    rowptr = np.zeros(lc.n + 1, np.int32); col = np.zeros(lc.col.size, np.int32); val = np.zeros(lc.col.size * lc.bs * lc.bs)
    ctx.call("uggpu_mat_get", k - 1, be.A, capi._p(rowptr), capi._p(col), capi._p(val))
    assert np.array_equal(rowptr, rp2) and np.array_equal(col, col2)
    # same values as on the reference's pattern, entry by entry (the order of a row's entries does not enter an entry's sum)
    bb = lc.bs * lc.bs
    ref = port.galerkin(k, golden.levels[k].val if k == namg else None).reshape(-1, bb)
    want = {}
    for r in range(lc.n):
        for e in range(lc.rowptr[r], lc.rowptr[r + 1]):
            want[(r, int(lc.col[e]))] = ref[e]
    got = val.reshape(-1, bb)
    for r in range(lc.n):
        for e in range(rowptr[r], rowptr[r + 1]):
            assert np.array_equal(got[e], want[(r, int(col[e]))]), (r, int(col[e]))
    be.close()


def test_gpu_amg_setup(golden):
    """uggpu_amg_coarsen_rs / uggpu_amg_coarsen_vanek rebuild every algebraic level of the Ruge-Stueben / Vanek dumps from the level above
    it -- strong connections, coarsening and interpolation, Galerkin matrix with the pattern created by the product -- and must arrive at
    the levels the reference's selectionAMG / clusterAMG built: flags, transfer stencils (list order, weights), matrix pattern (list
    order) and values, all bit for bit.  The solve records of the dump are then replayed on the rebuilt hierarchy."""
    from test_oracle_port import amg_levels, is_rs, vanek_config
    vk = vanek_config(golden)
    if not is_rs(golden) and vk is None:
        pytest.skip("dump without Ruge-Stueben or Vanek levels")
    import ctypes as C
    from backends import GpuBackend
    be = GpuBackend(golden)
    ctx = be.ctx
    namg = amg_levels(golden)
    for k in range(namg, 0, -1):
        nc = C.c_int(0)
        if vk is None:
            ctx.call("uggpu_amg_coarsen_rs", k, be.A, C.c_double(0.25), C.byref(nc))
        else:
            ctx.call("uggpu_amg_coarsen_vanek", k, be.A, C.c_double(vk[0]), int(vk[1]), C.byref(nc))
        assert nc.value == golden.levels[k - 1].n, k
    back = ctx.download_hierarchy(golden.top)
    for l in range(namg + 1):
        a, b = golden.levels[l], back.levels[l]
        for key in ("rowptr", "col", "val", "vclass", "vnclass", "ctl", "skip"):
            if l < namg or key in ("rowptr", "col", "val"):
                assert np.array_equal(getattr(a, key), getattr(b, key)), (l, key)
        if 0 < l <= namg:
            for key in ("p_rowptr", "p_col", "p_w", "r_rowptr", "r_col", "r_w"):
                assert np.array_equal(getattr(a, key), getattr(b, key)), (l, key)
    n = replay_solve(be, golden, exact=True, red_tol=1e-12)
    assert n > 10
    # below the last level the coarsening stops by itself one day: all or no vectors coarse -> no level, n_coarse = 0
    be.close()
