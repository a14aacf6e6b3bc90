"""oracle/ugport.py -- TEST INFRASTRUCTURE ONLY (ctypes view of oracle/libugport.so).

`PortBackend` exposes the CPU restatement (oracle/ugport.c) through the same small interface the
GPU test backend implements (tests/backends.py), so one replay of the reference's call sequence
(tests/replay.py) checks both against the golden dumps.  Nothing outside tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAX_BS = 3


class _Level(C.Structure):
    _fields_ = [("n", C.c_int), ("bs", C.c_int),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("vclass", C.c_void_p), ("vnclass", C.c_void_p), ("ctl", C.c_void_p), ("skip", C.c_void_p),
                ("p_rowptr", C.c_void_p), ("p_col", C.c_void_p), ("p_w", C.c_void_p),
                ("r_rowptr", C.c_void_p), ("r_col", C.c_void_p), ("r_w", C.c_void_p), ("ilu", C.c_void_p)]


class _Cfg(C.Structure):
    _fields_ = [("nu1", C.c_int), ("nu2", C.c_int), ("gamma", C.c_int), ("baselevel", C.c_int),
                ("smooth_damp", C.c_double * MAX_BS), ("cycle_damp", C.c_double * MAX_BS),
                ("base_maxit", C.c_int), ("base_reduction", C.c_double), ("base_abslimit", C.c_double),
                ("smoother", C.c_int), ("imat", C.c_int), ("imat_below", C.c_int), ("level_opt", C.c_int),
                ("base_hook", C.c_void_p), ("base_user", C.c_void_p)]


class _Fe(C.Structure):
    _fields_ = [("problem", C.c_int), ("dim", C.c_int), ("E", C.c_double), ("nu", C.c_double), ("source", C.c_double * MAX_BS)]


SMOOTHERS = {"jac": 0, "gs": 1, "sgs": 2, "sor": 3, "ilu": 4}


def build() -> str:
    path = os.path.join(_HERE, "libugport.so")
    src = os.path.join(_HERE, "ugport.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    return path


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ugport_base_factor.restype = C.c_void_p
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_ROWMODE = {"all": 0, "new_defect": 1, "fine_grid_dof": 2}


class PortBackend:
    """CPU restatement behind the test-backend interface."""

    name = "port"

    def __init__(self, hier):
        self.h = hier
        self.L = lib()
        self.bs = hier.bs
        self._keep = []
        arr = (_Level * len(hier.levels))()
        for i, lv in enumerate(hier.levels):
            fields = {}
            for k in ("rowptr", "col", "val", "vclass", "vnclass", "ctl", "skip",
                      "p_rowptr", "p_col", "p_w", "r_rowptr", "r_col", "r_w"):
                a = getattr(lv, k)
                if a is not None:
                    a = np.ascontiguousarray(a)
                    self._keep.append(a)
                fields[k] = _p(a)
            arr[i] = _Level(lv.n, lv.bs, ilu=None, **fields)
        self.levels = arr
        self.imat = bool(int(hier.raw["transfer_mode"][0])) if "transfer_mode" in getattr(hier, "raw", {}) else False
        # dumps with algebraic levels (ug_driver --amg): UG's levels < 1 always use the by-matrix transfer; the dump numbers them from 0
        self.imat_below = -int(hier.raw["bottomlevel"][0]) if "bottomlevel" in getattr(hier, "raw", {}) else 0
        self.vec: Dict[str, List[np.ndarray]] = {}
        self._ilu: Dict[int, tuple] = {}          # level -> (beta, decomposed values)

    # ---- vectors
    def _v(self, name, level):
        if name not in self.vec:
            self.vec[name] = [np.zeros(lv.n * lv.bs) for lv in self.h.levels]
        return self.vec[name][level]

    def put(self, level, name, a):
        self._v(name, level)[:] = a

    def get(self, level, name):
        return self._v(name, level).copy()

    def _lp(self, level):
        return C.byref(self.levels[level])

    def _surface(self, fl, tl, mode):
        """(level, rowmode) pairs of one reference loop (vecloop.ct / matloop.ct)."""
        if mode == 0:
            return [(l, 0) for l in range(fl, tl + 1)]
        return [(l, 2) for l in range(self.h.fullrefinelevel, tl)] + [(tl, 1)]

    def _vs(self, a):
        v = (C.c_double * MAX_BS)(*([0.0] * MAX_BS))
        for i, x in enumerate(np.atleast_1d(a)[:MAX_BS]):
            v[i] = float(x)
        return v

    # ---- BLAS 2
    def dmatmul(self, fl, tl, mode, op, x, y):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dmatmul(self._lp(l), op, rm, _dp(self._v(x, l)), _dp(self._v(y, l)))

    # ---- BLAS 1
    def dset(self, fl, tl, mode, x, a):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dset(self._lp(l), rm, _dp(self._v(x, l)), C.c_double(a))

    def dscal(self, fl, tl, mode, x, a):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dscal(self._lp(l), rm, _dp(self._v(x, l)), C.c_double(a))

    def dscalx(self, fl, tl, mode, x, a):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dscalx(self._lp(l), rm, _dp(self._v(x, l)), self._vs(a))

    def _xy(self, fn, fl, tl, mode, x, y):
        for l, rm in self._surface(fl, tl, mode):
            fn(self._lp(l), rm, _dp(self._v(x, l)), _dp(self._v(y, l)))

    def dcopy(self, fl, tl, mode, x, y): self._xy(self.L.ugport_dcopy, fl, tl, mode, x, y)
    def dadd(self, fl, tl, mode, x, y): self._xy(self.L.ugport_dadd, fl, tl, mode, x, y)
    def dsub(self, fl, tl, mode, x, y): self._xy(self.L.ugport_dsub, fl, tl, mode, x, y)
    def dminusadd(self, fl, tl, mode, x, y): self._xy(self.L.ugport_dminusadd, fl, tl, mode, x, y)

    def daxpy(self, fl, tl, mode, x, a, y):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_daxpy(self._lp(l), rm, _dp(self._v(x, l)), C.c_double(a), _dp(self._v(y, l)))

    def daxpyx(self, fl, tl, mode, x, a, y):
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_daxpyx(self._lp(l), rm, _dp(self._v(x, l)), self._vs(a), _dp(self._v(y, l)))

    def ddot(self, fl, tl, mode, x, y):
        s = C.c_double(0.0)
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_ddot_acc(self._lp(l), rm, _dp(self._v(x, l)), _dp(self._v(y, l)), C.byref(s))
        return s.value

    def ddotx(self, fl, tl, mode, x, y):
        s = self._vs([0.0])
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_ddotx_acc(self._lp(l), rm, _dp(self._v(x, l)), _dp(self._v(y, l)), s)
        return np.array(s[:self.bs])

    def dnrm2(self, fl, tl, mode, x):
        s = C.c_double(0.0)
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dnrm2_acc(self._lp(l), rm, _dp(self._v(x, l)), C.byref(s))
        return float(np.sqrt(s.value))

    def dnrm2x(self, fl, tl, mode, x):
        s = self._vs([0.0])
        for l, rm in self._surface(fl, tl, mode):
            self.L.ugport_dnrm2x_acc(self._lp(l), rm, _dp(self._v(x, l)), s)
        return np.sqrt(np.array(s[:self.bs]))

    # ---- smoother / transfer
    def l_jac(self, level, v, d):
        return self.L.ugport_l_jac(self._lp(level), _dp(self._v(v, level)), _dp(self._v(d, level)))

    def jac_smooth(self, level, x, b, damp):
        return self.L.ugport_jac_smooth(self._lp(level), _dp(self._v(x, level)), _dp(self._v(b, level)), self._vs(damp))

    def l_gs(self, level, v, d, upper=False, omega=None):
        """l_lgs / l_ugs (omega None) or l_lsor / l_usor."""
        name = "ugport_l_" + ("u" if upper else "l") + ("gs" if omega is None else "sor")
        args = [self._lp(level), _dp(self._v(v, level)), _dp(self._v(d, level))]
        if omega is not None:
            args.append(self._vs(omega))
        return getattr(self.L, name)(*args)

    def ilu_decomp(self, level, beta):
        """ILUPreProcess iter.cc:5444: L = copy of A, l_ilubthdecomp(L, beta); the level keeps the decomposition."""
        lv = self.h.levels[level]
        val = np.zeros(len(lv.col) * lv.bs * lv.bs)
        err = self.L.ugport_ilu_decomp(self._lp(level), self._vs([beta] * MAX_BS), _dp(val))
        if err == 0:
            self._ilu[level] = (float(beta), val)
            self.levels[level].ilu = _p(val)
        return err

    def ilu_values(self, level):
        return self._ilu[level][1].copy()

    def l_luiter(self, level, v, d):
        return self.L.ugport_l_luiter(self._lp(level), _dp(self._ilu[level][1]), _dp(self._v(v, level)), _dp(self._v(d, level)))

    def galerkin(self, level, fine_val):
        """AssembleGalerkinByMatrix on GRID_ON_LEVEL(level): values of the Galerkin matrix of level-1 (its own pattern) from the
        fine values `fine_val` (pattern of `level`)."""
        lc = self.h.levels[level - 1]
        fv = np.ascontiguousarray(fine_val, dtype=np.float64)
        out = np.zeros(len(lc.col) * lc.bs * lc.bs)
        err = self.L.ugport_galerkin(self._lp(level), self._lp(level - 1), _dp(fv), _dp(out))
        assert err == 0, err
        return out

    def galerkin_pattern(self, level, start=None):
        """Pattern (rowptr, col) of level-1 after AssembleGalerkinByMatrix on `level` had to create its connections; start = (rowptr, col)
        of the coarse level before the product, None = diagonal entries only."""
        lf, lc = self.h.levels[level], self.h.levels[level - 1]
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        ar, ac, pr, pc = i32(lf.rowptr), i32(lf.col), i32(lf.p_rowptr), i32(lf.p_col)
        sr, sc = (i32(start[0]), i32(start[1])) if start is not None else (None, None)
        rp = np.zeros(lc.n + 1, np.int32)
        args = [int(lf.n), int(lc.n), _p(ar), _p(ac), _p(pr), _p(pc), _p(sr), _p(sc)]
        assert self.L.ugport_galerkin_pattern(*args, _p(rp), None) == 0
        col = np.zeros(int(rp[-1]), np.int32)
        assert self.L.ugport_galerkin_pattern(*args, _p(rp), _p(col)) == 0
        return rp, col

    def assemble(self, level, fe, elem_ptr, elem_row, coef, coord, skip, x):
        """One level of LocalAssemble + AssembleDirichletBoundary (assemble.cc:657, disctools.cc:1837): returns (val, b)."""
        lv = self.h.levels[level]
        cfg = _Fe(fe["problem"], fe["dim"], fe["E"], fe["nu"], (C.c_double * MAX_BS)(*(list(fe["source"]) + [0.0] * MAX_BS)[:MAX_BS]))
        ep = np.ascontiguousarray(elem_ptr, dtype=np.int64); er = np.ascontiguousarray(elem_row, dtype=np.int32)
        cf = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        xy = np.ascontiguousarray(coord, dtype=np.float64); sk = np.ascontiguousarray(skip, dtype=np.uint32)
        xx = np.ascontiguousarray(x, dtype=np.float64)
        val = np.zeros(len(lv.col) * lv.bs * lv.bs); b = np.zeros(lv.n * lv.bs)
        self.L.ugport_assemble.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 8
        err = self.L.ugport_assemble(C.addressof(self.levels[level]), C.addressof(cfg), len(ep) - 1, _p(ep), _p(er), _p(cf), _p(xy), _p(sk), _p(xx), _p(val), _p(b))
        assert err == 0, err
        return val, b

    def smooth(self, level, kind, x, b, damp, tmp="__sgs"):
        return self.L.ugport_smooth(self._lp(level), SMOOTHERS[kind], _dp(self._v(x, level)), _dp(self._v(b, level)),
                                    self._vs(damp), _dp(self._v(tmp, level)))

    def restrict(self, level, to, frm, damp):
        (self.L.ugport_restrict_imat if self.imat or level <= self.imat_below else self.L.ugport_restrict)(self._lp(level), self._lp(level - 1), _dp(self._v(to, level - 1)),
                               _dp(self._v(frm, level)), self._vs(damp))

    def interpolate(self, level, to, frm, damp):
        (self.L.ugport_interpolate_imat if self.imat or level <= self.imat_below else self.L.ugport_interpolate)(self._lp(level), self._lp(level - 1), _dp(self._v(to, level)),
                                  _dp(self._v(frm, level - 1)), self._vs(damp))

    # ---- cycle / solver
    def _cfg(self, cfg):
        c = _Cfg()
        c.nu1, c.nu2, c.gamma, c.baselevel = cfg["nu1"], cfg["nu2"], cfg["gamma"], cfg.get("baselevel", 0)
        for i in range(MAX_BS):
            c.smooth_damp[i] = cfg["smooth_damp"]
            c.cycle_damp[i] = cfg.get("cycle_damp", 1.0)
        c.base_maxit = cfg.get("base_maxit", 10)
        c.base_reduction = cfg.get("base_reduction", 1e-8)
        c.base_abslimit = cfg.get("base_abslimit", 1e-10)
        c.smoother = SMOOTHERS[cfg.get("smoother", "jac")]
        c.imat = 1 if self.imat else 0
        c.imat_below = self.imat_below
        c.level_opt = int(cfg.get("level_opt", 0))
        if cfg.get("smoother") == "ilu":           # what LmgcPreProcess -> ILUPreProcess does on the levels above the base level
            beta = float(cfg.get("ilu_beta", 0.0))
            for l in range(c.baselevel + 1, len(self.h.levels)):
                if l not in self._ilu or self._ilu[l][0] != beta:
                    assert self.ilu_decomp(l, beta) == 0
        return c

    def _pp(self, name):
        n = len(self.h.levels)
        arr = (C.POINTER(C.c_double) * n)()
        for l in range(n):
            arr[l] = _dp(self._v(name, l))
        return arr

    def lmgc(self, level, c, b, cfg, t="__t"):
        cc = self._cfg(cfg)
        lu = C.c_void_p(self.L.ugport_base_factor(self._lp(cc.baselevel)))
        err = self.L.ugport_lmgc(self.levels, C.byref(cc), lu, level, self._pp(c), self._pp(b), self._pp(t))
        self.L.ugport_base_free(lu)
        return err

    def ls_defect(self, bl, level, x, b):
        self.L.ugport_ls_defect(self.levels, self.h.fullrefinelevel, bl, level, self._pp(x), self._pp(b))

    def ls_residuum(self, bl, level, b):
        d = self._vs([0.0])
        self.L.ugport_ls_residuum(self.levels, self.h.fullrefinelevel, bl, level, self._pp(b), d)
        return np.array(d[:self.bs])

    def solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, c="__c", t="__t"):
        cc = self._cfg(cfg)
        first = self._vs([0.0])
        hist = np.zeros(maxiter * self.bs)
        its = self.L.ugport_solve(self.levels, C.byref(cc), self.h.fullrefinelevel, level, self._pp(x), self._pp(b),
                                  self._pp(c), self._pp(t), maxiter, self._vs([abslimit] * MAX_BS),
                                  self._vs([reduction] * MAX_BS), first, _dp(hist))
        return its, np.array(first[:self.bs]), hist[:max(its, 0) * self.bs]

    def cg_solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, c="__c", t="__t", p="__p", tt="__tt"):
        cc = self._cfg(cfg)
        first = self._vs([0.0])
        hist = np.zeros(maxiter * self.bs)
        its = self.L.ugport_cg_solve(self.levels, C.byref(cc), self.h.fullrefinelevel, level, self._pp(x), self._pp(b), self._pp(c),
                                     self._pp(t), self._pp(p), self._pp(tt), maxiter, self._vs([abslimit] * MAX_BS),
                                     self._vs([reduction] * MAX_BS), first, _dp(hist))
        return its, np.array(first[:self.bs]), hist[:max(its, 0) * self.bs]

    def bcgs_solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, t="__t", weight=None):
        cc = self._cfg(cfg)
        first = self._vs([0.0])
        hist = np.zeros(maxiter * self.bs)
        w = self._vs([1.0] * MAX_BS if weight is None else [float(v) ** 2 for v in weight])
        work = [self._pp("__bcgs_" + n) for n in "rpvstq"]
        its = self.L.ugport_bcgs_solve(self.levels, C.byref(cc), self.h.fullrefinelevel, level, self._pp(x), self._pp(b), self._pp(t),
                                       *work, w, maxiter, self._vs([abslimit] * MAX_BS), self._vs([reduction] * MAX_BS), first, _dp(hist))
        return its, np.array(first[:self.bs]), hist[:((max(its, 0) + 1) // 2) * self.bs]
