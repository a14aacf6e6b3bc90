"""GpuBackend: the CUDA path behind the interface tests/replay.py drives (same as oracle.ugport.PortBackend).

Every method is one call through the C-ABI of include/uggpu.h -- the calls the `gpuls` numprocs make.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ug_b200 import capi


class GpuBackend:
    name = "gpu"

    def __init__(self, hier, fused=1, device=0):
        self.h = hier
        self.bs = hier.bs
        self.ctx = capi.Context(device)
        self.ctx.upload_hierarchy(hier, "A")
        self.A = self.ctx.handle("A")
        self.fused = fused
        self._have = set()

    def close(self):
        self.ctx.close()

    def _v(self, name, levels=None):
        """handle of `name`, allocated on every level (like a VECDATA_DESC)"""
        if name not in self._have:
            for l in range(len(self.h.levels)):
                self.ctx.alloc(l, name)
            self._have.add(name)
        return self.ctx.handle(name)

    def put(self, level, name, a):
        self._v(name)
        self.ctx.put(level, name, a)

    def get(self, level, name):
        self._v(name)
        return self.ctx.get(level, name)

    # ---- BLAS 2
    def dmatmul(self, fl, tl, mode, op, x, y):
        fn = ("uggpu_dmatmul", "uggpu_dmatmul_add", "uggpu_dmatmul_minus")[op]
        self.ctx.call(fn, fl, tl, mode, self._v(x), self.A, self._v(y))

    # ---- BLAS 1
    def dset(self, fl, tl, mode, x, a): self.ctx.call("uggpu_dset", fl, tl, mode, self._v(x), float(a))
    def dscal(self, fl, tl, mode, x, a): self.ctx.call("uggpu_dscal", fl, tl, mode, self._v(x), float(a))
    def dscalx(self, fl, tl, mode, x, a): self.ctx.call("uggpu_dscalx", fl, tl, mode, self._v(x), capi._vs(a))
    def dcopy(self, fl, tl, mode, x, y): self.ctx.call("uggpu_dcopy", fl, tl, mode, self._v(x), self._v(y))
    def dadd(self, fl, tl, mode, x, y): self.ctx.call("uggpu_dadd", fl, tl, mode, self._v(x), self._v(y))
    def dsub(self, fl, tl, mode, x, y): self.ctx.call("uggpu_dsub", fl, tl, mode, self._v(x), self._v(y))
    def dminusadd(self, fl, tl, mode, x, y): self.ctx.call("uggpu_dminusadd", fl, tl, mode, self._v(x), self._v(y))
    def daxpy(self, fl, tl, mode, x, a, y): self.ctx.call("uggpu_daxpy", fl, tl, mode, self._v(x), float(a), self._v(y))
    def daxpyx(self, fl, tl, mode, x, a, y): self.ctx.call("uggpu_daxpyx", fl, tl, mode, self._v(x), capi._vs(a), self._v(y))

    def ddot(self, fl, tl, mode, x, y):
        s = C.c_double(0.0)
        self.ctx.call("uggpu_ddot", fl, tl, mode, self._v(x), self._v(y), C.byref(s))
        return s.value

    def ddotx(self, fl, tl, mode, x, y):
        s = capi._vs([0.0])
        self.ctx.call("uggpu_ddotx", fl, tl, mode, self._v(x), self._v(y), s)
        return np.array(s[:self.bs])

    def dnrm2(self, fl, tl, mode, x):
        s = C.c_double(0.0)
        self.ctx.call("uggpu_dnrm2", fl, tl, mode, self._v(x), C.byref(s))
        return s.value

    def dnrm2x(self, fl, tl, mode, x):
        s = capi._vs([0.0])
        self.ctx.call("uggpu_dnrm2x", fl, tl, mode, self._v(x), s)
        return np.array(s[:self.bs])

    # ---- smoother / transfer
    def l_jac(self, level, v, d):
        return self.ctx.L.uggpu_l_jac(self.ctx.h, level, self._v(v), self.A, self._v(d))

    def jac_smooth(self, level, x, b, damp):
        return self.ctx.L.uggpu_jac_smooth(self.ctx.h, level, self._v(x), self._v(b), self.A, capi._vs(damp))

    def l_gs(self, level, v, d, upper=False, omega=None):
        """uggpu_l_lgs / uggpu_l_ugs (omega None) or uggpu_l_lsor / uggpu_l_usor."""
        fn = getattr(self.ctx.L, "uggpu_l_" + ("u" if upper else "l") + ("gs" if omega is None else "sor"))
        args = [self.ctx.h, level, self._v(v), self.A, self._v(d)]
        if omega is not None:
            args.append(capi._vs(omega))
        return fn(*args)

    # ---- ILU (uggpu_dmatcopy + uggpu_l_ilubthdecomp = ILUPreProcess iter.cc:5444, uggpu_l_luiter)
    def ilu_decomp(self, level, beta, L="__L"):
        self.ctx.call("uggpu_dmatcopy", level, level, 0, self.ctx.handle(L), self.A)
        return self.ctx.L.uggpu_l_ilubthdecomp(self.ctx.h, level, self.ctx.handle(L), capi._vs([beta]))

    def ilu_values(self, level, L="__L"):
        lv = self.h.levels[level]
        rowptr = np.zeros(lv.n + 1, np.int32); col = np.zeros(lv.col.size, np.int32); val = np.zeros(lv.col.size * lv.bs * lv.bs)
        self.ctx.call("uggpu_mat_get", level, self.ctx.handle(L), capi._p(rowptr), capi._p(col), capi._p(val))
        assert np.array_equal(rowptr, lv.rowptr) and np.array_equal(col, lv.col), "pattern of the decomposition differs from A"
        return val

    def l_luiter(self, level, v, d, L="__L"):
        return self.ctx.L.uggpu_l_luiter(self.ctx.h, level, self._v(v), self.ctx.handle(L), self._v(d))

    def galerkin(self, level, fine_val=None):
        """uggpu_galerkin: A of level-1 := P^T A_level P (AssembleGalerkinByMatrix after dmatset 0); returns its values.  fine_val is
        ignored: the device already holds the fine matrix (the cascade leaves the Galerkin matrix of the level above there)."""
        self.ctx.call("uggpu_galerkin", level, self.A)
        lc = self.h.levels[level - 1]
        rowptr = np.zeros(lc.n + 1, np.int32); col = np.zeros(lc.col.size, np.int32); val = np.zeros(lc.col.size * lc.bs * lc.bs)
        self.ctx.call("uggpu_mat_get", level - 1, self.A, capi._p(rowptr), capi._p(col), capi._p(val))
        assert np.array_equal(rowptr, lc.rowptr) and np.array_equal(col, lc.col)
        return val

    def smooth(self, level, kind, x, b, damp, tmp="__sgs"):
        t = self.ctx.handle("__L") if kind == "ilu" else self._v(tmp)        # ilu: the decomposition made by ilu_decomp
        return self.ctx.L.uggpu_smooth(self.ctx.h, level, capi.SMOOTHERS[kind], self._v(x), self._v(b), self.A, capi._vs(damp), t)

    def restrict(self, level, to, frm, damp):
        self.ctx.call("uggpu_restrict", level, self._v(to), self._v(frm), capi._vs(damp))

    def interpolate(self, level, to, frm, damp):
        self.ctx.call("uggpu_interpolate_correction", level, self._v(to), self._v(frm), capi._vs(damp))

    # ---- cycle / solver
    def _cfg(self, cfg, t="__t"):
        self._v(t)
        c = self.ctx.lmgc_cfg(nu1=cfg["nu1"], nu2=cfg["nu2"], gamma=cfg["gamma"], baselevel=cfg.get("baselevel", 0),
                              smooth_damp=cfg["smooth_damp"], cycle_damp=cfg.get("cycle_damp", 1.0),
                              base_maxit=cfg.get("base_maxit", 10), base_reduction=cfg.get("base_reduction", 1e-8),
                              base_abslimit=cfg.get("base_abslimit", 1e-10), fused=self.fused, t=t,
                              smoother=cfg.get("smoother", "jac"), ilu_beta=cfg.get("ilu_beta", 0.0), level_opt=cfg.get("level_opt", 0))
        return c

    def lmgc(self, level, c, b, cfg, t="__t"):
        cc = self._cfg(cfg, t)
        self.ctx.call("uggpu_lmgc_preprocess", C.byref(cc), level, self.A)
        return self.ctx.L.uggpu_lmgc(self.ctx.h, C.byref(cc), level, self._v(c), self._v(b), self.A)

    def ls_defect(self, bl, level, x, b):
        self.ctx.call("uggpu_ls_defect", bl, level, self._v(x), self._v(b), self.A)

    def ls_residuum(self, bl, level, b):
        r = capi.LResult()
        self.ctx.call("uggpu_ls_residuum", bl, level, self._v(b), C.byref(r))
        return np.array(r.last_defect[:self.bs])

    def solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, c="__c", t="__t"):
        cc = self._cfg(cfg, t)
        self.ctx.call("uggpu_lmgc_preprocess", C.byref(cc), level, self.A)
        r = capi.LResult()
        self.ctx.call("uggpu_ls_residuum", cc.baselevel, level, self._v(b), C.byref(r))
        hist = np.zeros(maxiter * self.bs)
        self.ctx.call("uggpu_ls_solve", C.byref(cc), cc.baselevel, level, self._v(x), self._v(b), self.A, self._v(c),
                      int(maxiter), capi._vs([abslimit]), capi._vs([reduction]), C.byref(r),
                      hist.ctypes.data_as(C.POINTER(C.c_double)))
        its = r.number_of_linear_iterations
        return its, np.array(r.first_defect[:self.bs]), hist[:its * self.bs]

    def cg_solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, c="__c", t="__t", p="__p", tt="__tt"):
        cc = self._cfg(cfg, t)
        self.ctx.call("uggpu_lmgc_preprocess", C.byref(cc), level, self.A)
        r = capi.LResult()
        self.ctx.call("uggpu_ls_residuum", cc.baselevel, level, self._v(b), C.byref(r))
        hist = np.zeros(maxiter * self.bs)
        self.ctx.call("uggpu_cg_solve", C.byref(cc), cc.baselevel, level, self._v(x), self._v(b), self.A, self._v(c), self._v(p),
                      self._v(tt), int(maxiter), capi._vs([abslimit]), capi._vs([reduction]), C.byref(r),
                      hist.ctypes.data_as(C.POINTER(C.c_double)))
        its = r.number_of_linear_iterations
        return its, np.array(r.first_defect[:self.bs]), hist[:its * self.bs]

    def bcgs_solve(self, level, x, b, cfg, maxiter, abslimit=1e-30, reduction=1e-30, t="__t", weight=None):
        cc = self._cfg(cfg, t)
        self.ctx.call("uggpu_lmgc_preprocess", C.byref(cc), level, self.A)
        r = capi.LResult()
        self.ctx.call("uggpu_ls_residuum", cc.baselevel, level, self._v(b), C.byref(r))
        hist = np.zeros(maxiter * self.bs)
        work = (C.c_int * 6)(*[self._v("__bcgs_" + n) for n in "rpvstq"])
        self.ctx.call("uggpu_bcgs_solve", C.byref(cc), cc.baselevel, level, self._v(x), self._v(b), self.A, work,
                      capi._vs([1.0] * capi.MAX_BS if weight is None else weight), 0, int(maxiter), capi._vs([abslimit]),
                      capi._vs([reduction]), C.byref(r), hist.ctypes.data_as(C.POINTER(C.c_double)))
        its = r.number_of_linear_iterations
        return its, np.array(r.first_defect[:self.bs]), hist[:((its + 1) // 2) * self.bs]
