"""Replays, on any backend, the exact sequence of reference calls that oracle/ug_driver.cc made
when it wrote a golden dump (dump_ops / dump_solve), and compares every result with the
reference's.  Backends: the CPU restatement (oracle/ugport.py, PortBackend) and the CUDA path
through the C-ABI (tests/backends.py, GpuBackend).

Bar (north_star): vectors bit-exact (`exact=True`; -0.0 == 0.0 counts as equal), reductions
(ddot/dnrm2 families) within `red_tol` relative because a parallel sum cannot follow the
reference's single running sum.
"""
from __future__ import annotations

import numpy as np

ALL = 0
SURF = -1


class Mismatch(AssertionError):
    pass


def has(hier, what):
    """Which record groups a dump holds: 'ops' (dump_ops), 'solve' (dump_solve), 'krylov' (dump_krylov); --lean dumps of
    oracle/ug_driver.cc leave groups out."""
    key = {"ops": "ops/a3", "solve": "solve/history", "krylov": "cg/K"}[what]
    return key in hier.raw


def _cmp_vec(what, got, ref, exact, tol, floor=0.0):
    """floor: lower bound of the scale of an inexact comparison (a defect that an exact base solve leaves at rounding level has
    no scale of its own)."""
    if got.shape != ref.shape:
        raise Mismatch(f"{what}: shape {got.shape} != {ref.shape}")
    if exact:
        if not np.array_equal(got, ref):
            bad = np.nonzero(got != ref)[0]
            i = bad[0]
            raise Mismatch(f"{what}: {bad.size}/{ref.size} entries differ, first at {i}: got {got[i]!r} ref {ref[i]!r}")
    else:
        scale = max(np.max(np.abs(ref)), floor) or 1.0
        err = np.max(np.abs(got - ref)) / scale
        if not err <= tol:
            raise Mismatch(f"{what}: max rel err {err:.3e} > {tol:.1e}")


def _cmp_red(what, got, ref, tol):
    got = np.atleast_1d(np.asarray(got, dtype=np.float64))
    ref = np.atleast_1d(ref)
    scale = np.maximum(np.abs(ref), 1e-300)
    err = np.max(np.abs(got - ref) / scale)
    if not err <= tol:
        raise Mismatch(f"{what}: rel err {err:.3e} > {tol:.1e} (got {got}, ref {ref})")


def replay_ops(be, hier, exact=True, vec_tol=1e-13, red_tol=1e-13, exact_red=False):
    """Mirror of dump_ops() in oracle/ug_driver.cc.  Returns the number of checks done."""
    d = hier.raw
    top = hier.top
    damp = d["damp"][0]
    a3 = d["ops/a3"]
    n = 0

    def chk(name, l, vec):
        nonlocal n
        _cmp_vec(f"L{l}/{name}", be.get(l, vec), d[f"L{l}/{name}"], exact, vec_tol)
        n += 1

    def red(name, got, ref):
        nonlocal n
        _cmp_red(name, got, ref, 0.0 if exact_red else red_tol)
        n += 1

    lean = "L0/dmatmul" not in d          # --lean dumps hold the smoother records only
    kind = smoother_of(hier)
    for l in range(top + 1):
        for v in "xbct":
            be.put(l, v, d[f"L{l}/in/{v}"])
        if f"L{l}/l_lgs" in d:            # Gauss-Seidel family
            for name, upper, omega in (("l_lgs", False, None), ("l_ugs", True, None), ("l_lsor", False, a3), ("l_usor", True, a3)):
                be.put(l, "t", d[f"L{l}/in/t"])
                assert be.l_gs(l, "t", "b", upper=upper, omega=omega) == 0
                chk(name, l, "t")
            be.put(l, "t", d[f"L{l}/in/t"])
        if f"L{l}/ilu/val" in d:          # ILU: decomposition values in canonical entry order, then l_luiter
            assert be.ilu_decomp(l, float(d["ilu_beta"][0])) == 0
            _cmp_vec(f"L{l}/ilu/val", be.ilu_values(l), d[f"L{l}/ilu/val"], exact, vec_tol); n += 1
            be.put(l, "t", d[f"L{l}/in/t"])
            assert be.l_luiter(l, "t", "b") == 0
            chk("l_luiter", l, "t")
            be.put(l, "t", d[f"L{l}/in/t"])
        if lean:
            be.put(l, "t", d[f"L{l}/in/t"])
            assert be.l_jac(l, "t", "b") == 0
            chk("l_jac", l, "t")
            if l > 0:
                be.put(l, "t", d[f"L{l}/in/t"])
                assert be.smooth(l, kind, "t", "b", [damp] * 3) == 0
                chk("smooth/t", l, "t"); chk("smooth/b", l, "b")
                be.put(l, "b", d[f"L{l}/in/b"])
            continue
        be.dmatmul(l, l, ALL, 0, "t", "x"); chk("dmatmul", l, "t")
        be.dmatmul(l, l, ALL, 1, "t", "b"); chk("dmatmul_add", l, "t")
        be.dmatmul(l, l, ALL, 2, "t", "c"); chk("dmatmul_minus", l, "t")
        be.put(l, "t", d[f"L{l}/in/t"])
        be.dcopy(l, l, ALL, "t", "x"); chk("dcopy", l, "t")
        be.dscal(l, l, ALL, "t", 0.75); chk("dscal", l, "t")
        be.dscalx(l, l, ALL, "t", a3); chk("dscalx", l, "t")
        be.dadd(l, l, ALL, "t", "b"); chk("dadd", l, "t")
        be.dsub(l, l, ALL, "t", "c"); chk("dsub", l, "t")
        be.dminusadd(l, l, ALL, "t", "b"); chk("dminusadd", l, "t")
        be.daxpy(l, l, ALL, "t", -1.375, "x"); chk("daxpy", l, "t")
        be.daxpyx(l, l, ALL, "t", a3, "c"); chk("daxpyx", l, "t")
        red(f"L{l}/ddot", be.ddot(l, l, ALL, "x", "b"), d[f"L{l}/ddot"])
        red(f"L{l}/dnrm2", be.dnrm2(l, l, ALL, "x"), d[f"L{l}/dnrm2"])
        red(f"L{l}/ddotx", be.ddotx(l, l, ALL, "x", "b"), d[f"L{l}/ddotx"])
        red(f"L{l}/dnrm2x", be.dnrm2x(l, l, ALL, "x"), d[f"L{l}/dnrm2x"])
        be.dset(l, l, ALL, "t", 0.5); chk("dset", l, "t")
        be.put(l, "t", d[f"L{l}/in/t"])
        assert be.l_jac(l, "t", "b") == 0
        chk("l_jac", l, "t")
        if l > 0:
            be.put(l, "t", d[f"L{l}/in/t"])
            assert be.smooth(l, kind, "t", "b", [damp] * 3) == 0
            chk("smooth/t", l, "t"); chk("smooth/b", l, "b")
            be.put(l, "b", d[f"L{l}/in/b"])
    if lean:
        return n

    for l in range(1, top + 1):
        be.put(l, "c", d[f"L{l}/restrict/in_fine"])
        be.put(l - 1, "c", d[f"L{l-1}/restrict/in_coarse"])
        be.restrict(l, "c", "c", a3)
        _cmp_vec(f"L{l}/restrict/out", be.get(l - 1, "c"), d[f"L{l-1}/restrict/out"], exact, vec_tol); n += 1
        be.put(l - 1, "x", d[f"L{l-1}/interpolate/in_coarse"])
        be.interpolate(l, "t", "x", a3)
        _cmp_vec(f"L{l}/interpolate/out", be.get(l, "t"), d[f"L{l}/interpolate/out"], exact, vec_tol); n += 1

    fr = hier.fullrefinelevel
    for l in range(top + 1):
        be.put(l, "x", d[f"L{l}/surf/in_x"]); be.put(l, "b", d[f"L{l}/surf/in_b"]); be.put(l, "t", d[f"L{l}/surf/in_t"])
    be.dmatmul(fr, top, SURF, 2, "b", "x")
    for l in range(top + 1):
        chk("surf/dmatmul_minus", l, "b")
    red("surf/dnrm2x", be.dnrm2x(fr, top, SURF, "b"), d["surf/dnrm2x"])
    red("surf/ddot", be.ddot(fr, top, SURF, "b", "x"), d["surf/ddot"])
    be.dset(fr, top, SURF, "t", 2.5)
    for l in range(top + 1):
        chk("surf/dset", l, "t")
    be.daxpy(fr, top, SURF, "t", 0.5, "x")
    for l in range(top + 1):
        chk("surf/daxpy", l, "t")
    return n


SMOOTHER_NAMES = ("jac", "gs", "sgs", "sor", "ilu")


def smoother_of(hier):
    """Smoother class the dump was written with (older dumps: jac)."""
    return SMOOTHER_NAMES[int(hier.raw["smoother"][0])] if "smoother" in hier.raw else "jac"


def cycle_cfg(hier, **over):
    d = hier.raw
    cfg = dict(nu1=int(d["nu1"][0]), nu2=int(d["nu2"][0]), gamma=int(d["gamma"][0]),
               baselevel=int(d["baselevel"][0]) if "baselevel" in d else 0, smoother=smoother_of(hier),
               smooth_damp=float(d["damp"][0]), cycle_damp=1.0, base_maxit=10, base_reduction=1e-8,
               base_abslimit=1e-10, ilu_beta=float(d["ilu_beta"][0]) if "ilu_beta" in d else 0.0,
               level_opt=int(d["level_opt"][0]) if "level_opt" in d else 0)
    cfg.update(over)
    return cfg


def replay_solve(be, hier, exact=True, vec_tol=1e-12, red_tol=1e-12, cfg_over=None):
    """Mirror of dump_solve(): one Lmgc cycle on the raw rhs, then `ls` runs of 1,2,5,N cycles."""
    d = hier.raw
    top = hier.top
    cfg = cycle_cfg(hier, **(cfg_over or {}))
    n = 0
    zeros = [np.zeros(lv.n * lv.bs) for lv in hier.levels]
    bfloor = float(np.max(np.abs(hier.levels[top].rhs)))       # scale of the defects
    for l in range(top + 1):
        be.put(l, "b", hier.levels[l].rhs); be.put(l, "c", zeros[l])
    assert be.lmgc(top, "c", "b", cfg) == 0
    for l in range(cfg["baselevel"], top + 1):                  # levels below the base level are not touched by the cycle
        _cmp_vec(f"L{l}/lmgc/c", be.get(l, "c"), d[f"L{l}/lmgc/c"], exact, vec_tol)
        _cmp_vec(f"L{l}/lmgc/b", be.get(l, "b"), d[f"L{l}/lmgc/b"], exact, vec_tol, bfloor)
        n += 2
    cycles = int(d["solve/cycles"][0])
    hist_ref = d["solve/history"].reshape(cycles, hier.bs)
    for k in sorted({1, 2, 5, cycles}):
        if k > cycles:
            continue
        for l in range(top + 1):
            be.put(l, "x", zeros[l]); be.put(l, "b", hier.levels[l].rhs)
        be.ls_defect(cfg["baselevel"], top, "x", "b")
        first = be.ls_residuum(cfg["baselevel"], top, "b")
        _cmp_red("solve/first_defect", first, d["solve/first_defect"], red_tol); n += 1
        its, first2, hist = be.solve(top, "x", "b", cfg, k)
        assert its == k, (its, k)
        _cmp_red("solve/first_defect(solver)", first2, d["solve/first_defect"], red_tol)
        _cmp_red(f"solve/history[:{k}]", hist.reshape(k, hier.bs), hist_ref[:k], red_tol); n += 2
        for l in range(cfg["baselevel"], top + 1):
            _cmp_vec(f"L{l}/solve/x_after_{k}", be.get(l, "x"), d[f"L{l}/solve/x_after_{k}"], exact, vec_tol)
            _cmp_vec(f"L{l}/solve/b_after_{k}", be.get(l, "b"), d[f"L{l}/solve/b_after_{k}"], exact, vec_tol, bfloor)
            n += 2
    return n


def replay_krylov(be, hier, exact=True, vec_tol=1e-10, red_tol=1e-9, floor=1e-10):
    """Mirror of dump_krylov(): class `cg` and class `bcgs` of the reference around the same cycle, fresh solves with
    $m 1, 2, K.  With exact=False (backends whose reductions are parallel sums) vectors are compared relative to their
    largest entry and history entries below floor * first_defect (rounding level of the converged solve) are skipped."""
    d = hier.raw
    top = hier.top
    cfg = cycle_cfg(hier)
    zeros = [np.zeros(lv.n * lv.bs) for lv in hier.levels]
    n = 0
    for name in ("cg", "bcgs"):
        K = int(d[f"{name}/K"][0])
        hist_ref = d[f"{name}/history"].reshape(K, hier.bs)
        its_ref = d[f"{name}/iterations"]
        first_ref = d[f"{name}/first_defect"]
        for k in sorted({1, 2, K}):
            for l in range(top + 1):
                be.put(l, "x", zeros[l]); be.put(l, "b", hier.levels[l].rhs)
            be.ls_defect(0, top, "x", "b")
            its, first, hist = (be.cg_solve if name == "cg" else be.bcgs_solve)(top, "x", "b", cfg, k)
            assert its == its_ref[k - 1], (name, k, its, its_ref[k - 1])
            _cmp_red(f"{name}/first_defect", first, first_ref, 0.0 if exact else 1e-12)
            got = hist.reshape(-1, hier.bs)[-1]
            ref = hist_ref[k - 1]
            if exact:
                _cmp_red(f"{name}/history[{k}]", got, ref, 0.0)
            else:
                keep = ref > floor * first_ref
                if keep.any():
                    _cmp_red(f"{name}/history[{k}]", got[keep], ref[keep], red_tol)
            n += 2
            for l in range(top + 1):
                _cmp_vec(f"L{l}/{name}/x_after_{k}", be.get(l, "x"), d[f"L{l}/{name}/x_after_{k}"], exact, vec_tol)
                if exact:
                    _cmp_vec(f"L{l}/{name}/b_after_{k}", be.get(l, "b"), d[f"L{l}/{name}/b_after_{k}"], True, vec_tol)
                else:       # the defect shrinks towards rounding level: compare on the scale of the first defect
                    gb, rb = be.get(l, "b"), d[f"L{l}/{name}/b_after_{k}"]
                    # (at least the first defect of the top level: the lower levels of a dump with algebraic levels start from zero)
                    scale = max(np.max(np.abs(d[f"L{l}/solve/b_first"])), np.max(np.abs(d[f"L{top}/solve/b_first"])), 1e-300)
                    if not np.max(np.abs(gb - rb)) <= vec_tol * scale:
                        raise Mismatch(f"L{l}/{name}/b_after_{k}: abs err {np.max(np.abs(gb - rb)):.3e} > {vec_tol:.1e} * {scale:.3e}")
                n += 2
    return n
