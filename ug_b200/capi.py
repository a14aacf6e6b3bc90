"""ctypes binding of the C-ABI in include/uggpu.h (ug_b200/lib/libuggpu.so).

This is the same boundary the C++ `gpuls` numprocs (ug_b200/host/gpuls_np.cc) call; the Python side exists for
the tests and the bench.  There is no fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("UGGPU_LIB") or os.path.join(ROOT, "ug_b200", "lib", "libuggpu.so")
HEADER = os.path.join(ROOT, "include", "uggpu.h")

MAX_BS = 3
ALL_VECTORS = 0
ON_SURFACE = -1
SYNTH_P1_SIMPLEX, SYNTH_Q1_POISSON, SYNTH_Q1_ELASTICITY, SYNTH_P1_VARCOEF = 0, 1, 2, 3


class UggpuError(RuntimeError):
    pass


BaseSolverFn = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int)


class LmgcCfg(C.Structure):
    _fields_ = [("nu1", C.c_int), ("nu2", C.c_int), ("gamma", C.c_int), ("baselevel", C.c_int),
                ("smooth_damp", C.c_double * MAX_BS), ("cycle_damp", C.c_double * MAX_BS),
                ("t", C.c_int), ("base_maxit", C.c_int), ("base_reduction", C.c_double), ("base_abslimit", C.c_double),
                ("base_solver", C.c_void_p), ("base_user", C.c_void_p), ("fused", C.c_int), ("smoother", C.c_int),
                ("smoother_L", C.c_int), ("ilu_beta", C.c_double * MAX_BS), ("level_opt", C.c_int)]


SMOOTHERS = {"jac": 0, "gs": 1, "sgs": 2, "sor": 3, "ilu": 4}       # UGGPU_SM_*


class FeCfg(C.Structure):               # uggpu_fe_cfg
    _fields_ = [("problem", C.c_int), ("dim", C.c_int), ("E", C.c_double), ("nu", C.c_double), ("source", C.c_double * MAX_BS)]


class LResult(C.Structure):
    _fields_ = [("error_code", C.c_int), ("converged", C.c_int), ("number_of_linear_iterations", C.c_int),
                ("first_defect", C.c_double * MAX_BS), ("last_defect", C.c_double * MAX_BS)]


def declared_symbols() -> List[str]:
    """Every function name include/uggpu.h declares."""
    with open(HEADER) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uggpu_[a-z0-9_]+)\s*\(", src)) - {"uggpu_base_solver_fn"})


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UggpuError(f"{LIB_PATH} is missing: build it with `make -C ug_b200/csrc` (or __graft_entry__.build()); "
                             "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.uggpu_last_error.restype = C.c_char_p
        for name in ("uggpu_launch_count", "uggpu_device_bytes", "uggpu_mat_nnz", "uggpu_mat_padded_nnz", "uggpu_transfer_nnz", "uggpu_mat_col_words", "uggpu_mat_val_entries", "uggpu_mat_stencil_slices"):
            getattr(L, name).restype = C.c_int64
        L.uggpu_mat_pass_bytes.restype = C.c_double
        L.uggpu_dset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.uggpu_dscal.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.uggpu_daxpy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _vs(a, n=MAX_BS):
    v = (C.c_double * MAX_BS)(*([0.0] * MAX_BS))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    for i in range(MAX_BS):
        v[i] = float(a[i] if i < a.size else a[-1])
    return v


class Context:
    """One uggpu context (one GPU, one stream, one multigrid hierarchy)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.uggpu_ctx_create(int(device), C.byref(h))
        if rc:
            raise UggpuError(f"uggpu_ctx_create({device}) -> {rc}: {self.L.uggpu_last_error().decode()}")
        self.h = h
        self._names = {}

    def close(self):
        if self.h:
            self.L.uggpu_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, fn: str, *args):
        rc = getattr(self.L, fn)(self.h, *args)
        if rc:
            raise UggpuError(f"{fn} -> {rc}: {self.L.uggpu_last_error().decode()}")
        return rc

    def call_noctx(self, fn: str, *args):
        rc = getattr(self.L, fn)(*args)
        if rc:
            raise UggpuError(f"{fn} -> {rc}: {self.L.uggpu_last_error().decode()}")
        return rc

    # ---- handles: descriptor names -> small ints (what VECDATA_DESC/MATDATA_DESC pointers are to the numprocs)
    def handle(self, name: str) -> int:
        if name not in self._names:
            self._names[name] = len(self._names) + 1
        return self._names[name]

    # ---- hierarchy
    def upload_hierarchy(self, hier, A: str = "A"):
        a = self.handle(A)
        for l, lv in enumerate(hier.levels):
            self.call("uggpu_level_create", l, int(lv.n), int(lv.bs))
            self.call("uggpu_level_set_flags", l, _p(np.ascontiguousarray(lv.vclass)), _p(np.ascontiguousarray(lv.vnclass)),
                      _p(np.ascontiguousarray(lv.ctl)), _p(np.ascontiguousarray(lv.skip)))
            self.call("uggpu_mat_set", l, a, _p(np.ascontiguousarray(lv.rowptr)), _p(np.ascontiguousarray(lv.col)),
                      _p(np.ascontiguousarray(lv.val)))
            if l > 0:
                self.call("uggpu_transfer_set", l, _p(np.ascontiguousarray(lv.p_rowptr)), _p(np.ascontiguousarray(lv.p_col)),
                          _p(np.ascontiguousarray(lv.p_w)), _p(np.ascontiguousarray(lv.r_rowptr)),
                          _p(np.ascontiguousarray(lv.r_col)), _p(np.ascontiguousarray(lv.r_w)))
                if f"L{l}/transfer_mode" in hier.raw:    # dumps with algebraic levels (--amg): by-matrix transfer below UG's level 1
                    self.call("uggpu_transfer_set_mode", l, int(hier.raw[f"L{l}/transfer_mode"][0]))
                elif "transfer_mode" in hier.raw:        # dumps written with `transfer $M` (oracle/ug_driver.cc --imat)
                    self.call("uggpu_transfer_set_mode", l, int(hier.raw["transfer_mode"][0]))
        self.call("uggpu_set_fullrefinelevel", int(hier.fullrefinelevel))

    def upload_local_levels(self, levels, fullrefinelevel: int, bs: int, A: str = "A"):
        """One rank's part of a partitioned hierarchy (ug_b200.partition.split): levels, partitions, flags, matrices, transfer stencils."""
        a = self.handle(A)
        i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        for l, lv in enumerate(levels):
            self.call("uggpu_level_create", l, int(lv.n), int(bs))
            if lv.partitioned:
                self.call("uggpu_level_set_partition", l, int(lv.n_ghost), C.c_int64(int(lv.n_global)), int(lv.nb_rank.size), _p(i32(lv.nb_rank)),
                          _p(i32(lv.send_off)), _p(i32(lv.send_idx)), _p(i32(lv.recv_off)))
            self.call("uggpu_level_set_flags", l, _p(np.ascontiguousarray(lv.vclass)), _p(np.ascontiguousarray(lv.vnclass)),
                      _p(np.ascontiguousarray(lv.ctl)), _p(np.ascontiguousarray(lv.skip)))
            self.call("uggpu_mat_set", l, a, _p(i32(lv.rowptr)), _p(i32(lv.col)), _p(np.ascontiguousarray(lv.val, dtype=np.float64)))
            if l > 0:
                self.call("uggpu_transfer_set", l, _p(i32(lv.p_rowptr)), _p(i32(lv.p_col)), _p(np.ascontiguousarray(lv.p_w, dtype=np.float64)),
                          _p(i32(lv.r_rowptr)), _p(i32(lv.r_col)), _p(np.ascontiguousarray(lv.r_w, dtype=np.float64)))
        self.call("uggpu_set_fullrefinelevel", int(fullrefinelevel))

    def download_hierarchy(self, top: int, A: str = "A"):
        """Canonical CSR/flags/stencils of levels 0..top back on the host (ug_b200.hierarchy.Hierarchy)."""
        from .hierarchy import Hierarchy, Level
        a = self.handle(A)
        levels = []
        for l in range(top + 1):
            n, bs = self.level_n(l), self.level_bs(l)
            nnz = int(self.L.uggpu_mat_nnz(self.h, l, a))
            rowptr = np.zeros(n + 1, np.int32); col = np.zeros(nnz, np.int32); val = np.zeros(nnz * bs * bs)
            self.call("uggpu_mat_get", l, a, _p(rowptr), _p(col), _p(val))
            vclass = np.zeros(n, np.uint8); vnclass = np.zeros(n, np.uint8); ctl = np.zeros(n, np.uint8); skip = np.zeros(n, np.uint32)
            self.call("uggpu_level_get_flags", l, _p(vclass), _p(vnclass), _p(ctl), _p(skip))
            lv = Level(n=n, bs=bs, rowptr=rowptr, col=col, val=val, vclass=vclass, vnclass=vnclass, ctl=ctl, skip=skip)
            if l > 0:
                for which, pre in ((0, "p"), (1, "r")):
                    nrows = n if which == 0 else levels[l - 1].n
                    z = int(self.L.uggpu_transfer_nnz(self.h, l, which))
                    rp = np.zeros(nrows + 1, np.int32); cc = np.zeros(z, np.int32); ww = np.zeros(z)
                    self.call("uggpu_transfer_get", l, which, _p(rp), _p(cc), _p(ww))
                    setattr(lv, pre + "_rowptr", rp); setattr(lv, pre + "_col", cc); setattr(lv, pre + "_w", ww)
            levels.append(lv)
        return Hierarchy(dim=0, bs=levels[0].bs, fullrefinelevel=top, levels=levels)

    def level_n(self, l): return int(self.L.uggpu_level_n(self.h, l))
    def level_bs(self, l): return int(self.L.uggpu_level_bs(self.h, l))

    # ---- vectors
    def put(self, level: int, name: str, a: np.ndarray):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.level_n(level) * self.level_bs(level), (a.size, level)
        self.call("uggpu_vec_upload", level, self.handle(name), _p(a))

    def get(self, level: int, name: str) -> np.ndarray:
        out = np.empty(self.level_n(level) * self.level_bs(level))
        self.call("uggpu_vec_download", level, self.handle(name), _p(out))
        return out

    def alloc(self, level: int, name: str):
        self.call("uggpu_vec_alloc", level, self.handle(name))

    def devptr(self, level: int, name: str) -> int:
        p = C.c_void_p()
        self.call("uggpu_vec_devptr", level, self.handle(name), C.byref(p))
        return p.value

    def stream(self) -> int:
        p = C.c_void_p()
        self.call("uggpu_stream", C.byref(p))
        return p.value or 0

    def halo_transport(self) -> str:
        return {0: "none", 1: "nccl send/recv", 2: "peer-memory windows (CUDA IPC, push + unpack kernels)",
                3: "peer-memory ghost rows (CUDA IPC, pushes fused into the producing kernels)"}.get(int(self.L.uggpu_comm_transport(self.h)), "?")

    def assemble(self, level: int, x: str, b: str, A: str, fe: dict, elem_ptr, elem_row, coef, coord, skip):
        """uggpu_assemble: one level of LocalAssemble + AssembleDirichletBoundary on the device (fe: problem, dim, E, nu, source)."""
        cfg = FeCfg(int(fe["problem"]), int(fe["dim"]), float(fe.get("E", 1.0)), float(fe.get("nu", 0.3)),
                    (C.c_double * MAX_BS)(*(list(fe["source"]) + [0.0] * MAX_BS)[:MAX_BS]))
        ep = np.ascontiguousarray(elem_ptr, dtype=np.int64); er = np.ascontiguousarray(elem_row, dtype=np.int32)
        cf = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
        xy = np.ascontiguousarray(coord, dtype=np.float64)
        sk = None if skip is None else np.ascontiguousarray(skip, dtype=np.uint32)
        self.call("uggpu_assemble", level, self.handle(x), self.handle(b), self.handle(A), C.byref(cfg), C.c_int64(len(ep) - 1), _p(ep), _p(er), _p(cf), _p(xy), _p(sk))

    def mat_values(self, level: int, A: str, nnz: int) -> np.ndarray:
        bs = self.level_bs(level)
        val = np.zeros(nnz * bs * bs)
        self.call("uggpu_mat_get", level, self.handle(A), None, None, _p(val))
        return val

    def sync(self): self.call("uggpu_sync")
    def launch_count(self) -> int: return int(self.L.uggpu_launch_count(self.h))
    def device_bytes(self) -> int: return int(self.L.uggpu_device_bytes(self.h))

    # ---- cycle configuration
    def lmgc_cfg(self, nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=0.6, cycle_damp=1.0, base_maxit=10,
                 base_reduction=1e-8, base_abslimit=1e-10, fused=1, t="__t", smoother="jac", ilu_beta=0.0, smoother_L="__L", level_opt=0) -> LmgcCfg:
        c = LmgcCfg()
        c.nu1, c.nu2, c.gamma, c.baselevel = nu1, nu2, gamma, baselevel
        for i in range(MAX_BS):
            c.smooth_damp[i] = smooth_damp
            c.cycle_damp[i] = cycle_damp
        c.t = self.handle(t)
        c.base_maxit, c.base_reduction, c.base_abslimit = base_maxit, base_reduction, base_abslimit
        c.base_solver = None
        c.base_user = None
        c.fused = int(fused)
        c.smoother = SMOOTHERS[smoother]
        c.smoother_L = self.handle(smoother_L)
        c.level_opt = int(level_opt)
        for i in range(MAX_BS):
            c.ilu_beta[i] = ilu_beta
        return c
