#!/usr/bin/env python
"""bench.py -- V(2,2)-cycle throughput of the gpuls hot path on B200 (BASELINE.json metric).

One "step" = one iteration of the linear solver `ls` with `lmgc` V(2,2) damped-Jacobi as Iter, i.e. exactly the body
of LinearSolver's loop (np/procs/ls.cc:693-708): c = 0; c = Lmgc(b) (b updated to the new defect); x += c;
||b||_2 -- on a synthetic 3D P1 Poisson hierarchy (BASELINE.json configs[1]: unit cube, tetrahedra, base 4x4x4 cells,
7 uniform refinements = 8 levels, 513^3 = 135 005 697 fine unknowns).

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference [...]                        the unmodified reference on the host CPU (oracle/_ref)

Prints ONE JSON line (rank 0).  `value` = fine unknowns * K / device time of K steps with everything resident in HBM;
`e2e` = the same through the C-ABI with HOST vectors (x, b uploaded and downloaded inside the timed region, as
NP_LINEAR_SOLVER::Solver does with UG's VECTOR lists); `roofline` = the dominant kernel (fused smoothing step on the
finest level) timed with CUDA events around each of its launches inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vcycle_unknowns_per_s"
UNIT = "unknowns/s"


def workload_name(cells, top, n, kind="p1", smoother="jac"):
    what = {"p1": "3D P1 Poisson, unit cube, Kuhn tetrahedra", "q1": "3D Q1 Poisson, unit cube, hexahedra (27-point rows)",
            "elasticity": "3D Q1 linear elasticity (3x3 blocks, 27 block entries per row), unit cube, hexahedra"}[kind]
    return (f"{what}, base {cells}x{cells}x{cells} cells, {top + 1} levels, "
            f"{n} fine unknowns, V(2,2) " + {"jac": f"{'block-' if kind == 'elasticity' else ''}Jacobi damp 0.6", "gs": "Gauss-Seidel damp 1.0",
                                             "sgs": "symmetric Gauss-Seidel damp 1.0", "sor": "SOR omega 1.1", "ilu": "ILU(0) beta 0 damp 1.0"}[smoother] + ", base solver ls+lu")


# ---------------------------------------------------------------------------------------------------------------
def host_replicas(per_replica_gb: float = 1.0) -> int:
    """How many independent single-threaded UG replicas this host runs at once: one per core the process may use,
    bounded by memory (a 65^3 hierarchy needs ~0.7 GB in UG's data structures) and by 64."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    try:
        with open("/proc/meminfo") as f:
            avail_gb = next(int(l.split()[1]) for l in f if l.startswith("MemAvailable")) / 1e6
    except Exception:
        avail_gb = 8.0
    return max(1, min(cores, int(avail_gb * 0.6 / per_replica_gb), 64))


def run_reference_cpu(refine: int, cycles: int, replicas: int = 0):
    """Times the UNMODIFIED reference (oracle/_ref/ugoracle3: UG's own ls+lmgc+jac+transfer numprocs on its VECTOR/
    MATRIX lists) on this host.  UG is single-threaded (its only parallel mode is MPI, not installed), so all host cores
    are used the way an MPI run would use them at best: `replicas` independent copies of the same problem run
    concurrently (rendezvous after the hierarchy is built, oracle/ug_driver.cc --barrier) and their throughputs add up."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ugoracle3")
    if not os.path.exists(exe):
        return None
    replicas = replicas or host_replicas()
    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([exe, "--grid", "tet", "--refine", str(refine), "--damp", "0.6", "--time", "--cycles", str(cycles),
                                   "--reps", "1", "--barrier", d, str(replicas), str(i)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(replicas)]
        outs = [p.communicate(timeout=1500) for p in procs]
    res = []
    for so, se in outs:
        for line in so.splitlines():
            if line.startswith('{"kind"'):
                res.append(json.loads(line))
                break
        else:
            raise RuntimeError("reference run produced no result line:\n" + so[-2000:] + se[-2000:])
    r = dict(res[0])
    r["cores"] = replicas
    r["vcycle_unknowns_per_s"] = sum(x["vcycle_unknowns_per_s"] for x in res)
    r["s_per_cycle"] = max(x["s_per_cycle"] for x in res)
    r["per_core_unknowns_per_s"] = r["vcycle_unknowns_per_s"] / replicas
    return r


def run_port_cpu(cells: int, top: int, cycles: int):
    """Fallback CPU baseline when oracle/_ref is absent: the plain-C restatement (oracle/ugport.c) on a synthetic
    hierarchy downloaded from the device generator."""
    import numpy as np
    from ug_b200 import capi
    from oracle.ugport import PortBackend
    ctx = capi.Context(0)
    ctx.call("uggpu_synth_hierarchy", capi.SYNTH_P1_SIMPLEX, cells, cells, cells, top, ctx.handle("A"))
    hier = ctx.download_hierarchy(top)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    rhs = ctx.get(top, "b")
    ctx.close()
    be = PortBackend(hier)
    for l, lv in enumerate(hier.levels):
        be.put(l, "x", np.zeros(lv.n)); be.put(l, "b", rhs if l == top else np.zeros(lv.n))
    cfg = dict(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=0.6)
    t0 = time.perf_counter()
    its, _, _ = be.solve(top, "x", "b", cfg, cycles)
    dt = time.perf_counter() - t0
    n = hier.levels[-1].n
    return {"kind": "port", "cores": 1, "levels": top + 1, "unknowns": n, "cycles": its, "s_per_cycle": dt / its, "vcycle_unknowns_per_s": n * its / dt}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cycles = max(1, min(args.steps, 20))
    r = run_reference_cpu(args.cpu_refine, cycles, args.cpu_replicas)
    kind = "reference"
    if r is None:
        r = run_port_cpu(1, 5, cycles)
        kind = "port"
    n = r["unknowns"]
    sample = (f"{'UG 3.12.1 ls+lmgc+jac+transfer' if kind == 'reference' else 'oracle/ugport.c'}: {r['cycles']} V(2,2) cycles on a "
              f"{r.get('levels', '?')}-level unit-cube tet hierarchy with {n} fine unknowns (largest the host builds in ~10 s), "
              f"{r['cores']} concurrent single-threaded replica(s), throughputs added")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["vcycle_unknowns_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["cycles"], "warmup": 0, "ms_per_step": r["s_per_cycle"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.cells, args.top, (args.cells * 2 ** args.top + 1) ** 3),
                   "measured_on": sample},
        "cpu_baseline": {"value": r["vcycle_unknowns_per_s"], "unit": UNIT, "cores": r["cores"], "kind": kind, "sample": sample},
        "e2e": {"value": r["vcycle_unknowns_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [s.strip() for s in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


def our_arm(args):
    import numpy as np
    import torch
    from ug_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gpuls path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cells, top = args.cells, args.top
    ctx = capi.Context(local)
    A = ctx.handle("A")
    t0 = time.perf_counter()
    # weak scaling: every GPU gets a box of cells^3 base cells (513^3-type fine grid per GPU); the rank array follows
    # UG's RCB of a structured grid (2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2)
    P = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world)
    if P is None:
        raise SystemExit(f"bench.py: unsupported GPU count {world} (1, 2, 4 or 8)")
    if world > 1:
        import torch.distributed as dist
        idbuf = (C.c_char * 128)()
        if rank == 0:
            ctx.call_noctx("uggpu_comm_unique_id", idbuf)
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        ctx.call("uggpu_comm_init", world, rank, C.c_char_p(bytes(t.cpu().tolist())))
    kind = {"p1": capi.SYNTH_P1_SIMPLEX, "q1": capi.SYNTH_Q1_POISSON, "elasticity": capi.SYNTH_Q1_ELASTICITY}[args.kind]
    ctx.call("uggpu_synth_hierarchy_part", kind, cells * P[0], cells * P[1], cells * P[2], top, A,
             P[0], P[1], P[2], rank, C.c_int64(args.replicate_below))
    bs = ctx.level_bs(top)
    n = ctx.level_n(top) * bs                      # unknowns (UG counts vector components), this rank
    n_global = int(ctx.L.uggpu_level_n_global(ctx.h, top)) * bs
    for name in ("x", "b", "c"):
        for l in range(top + 1):
            ctx.alloc(l, name)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    sm_damp = {"jac": 0.6, "gs": 1.0, "sgs": 1.0, "sor": 1.1, "ilu": 1.0}[args.smoother]
    cfg = ctx.lmgc_cfg(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=sm_damp, fused=1, smoother=args.smoother)
    galerkin_ms = None
    if args.galerkin:      # setup operation, outside the timed steps: A_{l-1} := P^T A_l P cascaded from the top level down (uggpu_galerkin)
        if world > 1:
            raise SystemExit("bench.py: --galerkin runs on one GPU")
        galerkin_ms = []
        for l in range(top, 0, -1):
            ctx.sync()
            tg = time.perf_counter()
            ctx.call("uggpu_galerkin", l, A)
            ctx.sync()
            galerkin_ms.append(round((time.perf_counter() - tg) * 1e3, 3))
    ctx.sync()
    t_pre = time.perf_counter()
    ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
    ctx.sync()
    preprocess_s = time.perf_counter() - t_pre      # LmgcPreProcess: base-level LU, Gauss-Seidel schedules / ILU decompositions of all levels
    setup_s = time.perf_counter() - t0
    X, B, Cc = ctx.handle("x"), ctx.handle("b"), ctx.handle("c")
    res = capi.LResult()
    ctx.call("uggpu_ls_defect", 0, top, X, B, A)
    ctx.call("uggpu_ls_residuum", 0, top, B, C.byref(res))
    first = res.last_defect[0]
    absl, red = capi._vs([1e-300]), capi._vs([1e-300])

    def step():
        ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, X, B, A, Cc, 1, absl, red, C.byref(res), None)

    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    hist = []
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ctx.call("uggpu_prof_enable", 1)
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
        hist.append(res.last_defect[0])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    # per-kernel event times of the timed region
    prof = {}
    kinds = {"smooth": 0, "jac": 1, "restrict": 2, "interpolate": 3, "vecop": 4, "reduce": 5, "dmatmul": 6, "base": 7, "trisolve": 8}
    cnt, kms, kby = C.c_int64(), C.c_double(), C.c_double()
    for name, k in kinds.items():
        ctx.call("uggpu_prof_summary", k, -1, C.byref(cnt), C.byref(kms), C.byref(kby))
        prof[name] = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
    # dominant kernel: the fused smoothing step (Jacobi) / the defect update dmatmul_minus (Gauss-Seidel family), finest level
    ctx.call("uggpu_prof_summary", 0 if args.smoother == "jac" else 6, top, C.byref(cnt), C.byref(kms), C.byref(kby))
    dom = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
    ctx.call("uggpu_prof_summary", 8, top, C.byref(cnt), C.byref(kms), C.byref(kby))
    tri = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
    ctx.call("uggpu_prof_enable", 0)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- end to end: host vectors in, host vectors out, every step -------------------------------------------------
    xh = torch.zeros(n, dtype=torch.float64).pin_memory()      # n counts components: the vectors hold n doubles
    bh = torch.empty(n, dtype=torch.float64).pin_memory()
    ctx.call("uggpu_vec_download", top, X, C.c_void_p(xh.data_ptr()))
    ctx.call("uggpu_vec_download", top, B, C.c_void_p(bh.data_ptr()))

    def e2e_step():
        # b first (the cycle starts with it); x follows on the copy stream and hides behind the cycle, whose last kernel
        # is the first to touch it
        ctx.call("uggpu_vec_upload", top, B, C.c_void_p(bh.data_ptr()))
        ctx.call("uggpu_vec_upload_async", top, X, C.c_void_p(xh.data_ptr()))
        ctx.call("uggpu_ls_residuum", 0, top, B, C.byref(res))
        step()
        ctx.call("uggpu_vec_download", top, X, C.c_void_p(xh.data_ptr()))
        ctx.call("uggpu_vec_download", top, B, C.c_void_p(bh.data_ptr()))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if sampler else None
    dev_bytes = ctx.device_bytes()
    # BASELINE.json's second figure, "SpMV HBM GB/s vs peak": the plain dmatmul_minus (c -= A x) on the finest level, timed by the
    # library's per-kernel events after everything else (c is work space: every cycle starts by overwriting it)
    spmv = None
    if world == 1:
        try:
            ctx.call("uggpu_dmatmul_minus", top, top, 0, Cc, A, X)
            ctx.sync()
            ctx.call("uggpu_prof_enable", 1)
            for _ in range(5):
                ctx.call("uggpu_dmatmul_minus", top, top, 0, Cc, A, X)
            ctx.sync()
            c2, m2, b2 = C.c_int64(), C.c_double(), C.c_double()
            ctx.call("uggpu_prof_summary", 6, top, C.byref(c2), C.byref(m2), C.byref(b2))
            ctx.call("uggpu_prof_enable", 0)
            if c2.value > 0 and m2.value > 0:
                spmv = {"kernel": f"k_dmatmul_k<{bs},2> (x -= A y, finest level)", "launches": int(c2.value), "avg_ms": m2.value / c2.value,
                        "alg_bytes_per_launch": b2.value / c2.value, "GBps": b2.value / (m2.value * 1e-3) / 1e9}
        except Exception as e:          # a reported figure, never a reason to lose the bench line
            spmv = {"error": str(e)[:200]}

    if rank != 0:
        return 0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):        # ncu capture of the default workload only
        tj = json.load(open(tpath))
        if tj.get("workload") == {"kind": args.kind, "cells": cells, "top": top} and world == 1:
            traffic = tj.get("dominant_kernel_dram_bytes_per_launch", tj.get("k_smooth_k_dram_bytes_per_launch"))
    # SURVEY.md 8(d) counts 4 bytes of column index per entry; the stored format fetches fewer (compressed column words)
    nnz_top, words_top = int(ctx.L.uggpu_mat_nnz(ctx.h, top, A)), int(ctx.L.uggpu_mat_col_words(ctx.h, top, A))
    vals_top = int(ctx.L.uggpu_mat_val_entries(ctx.h, top, A))      # entries whose values a sweep fetches (shared value tables)
    sten_top = int(ctx.L.uggpu_mat_stencil_slices(ctx.h, top, A)) if not os.environ.get("UGGPU_NO_STENCIL") else 0
    sten_w = round(nnz_top / max(ctx.level_n(top), 1))
    if sten_top > 0 and bs == 1 and sten_w in (15, 27):
        smooth_kernel = f"k_smooth_sten<*,{sten_w}> (fused smoothing step, stencil variant, finest level)"
    elif sten_top > 0 and bs == 3:
        smooth_kernel = "k_smooth_sten3<*> (fused smoothing step, 3x3-block stencil variant, finest level)"
    else:
        smooth_kernel = f"k_smooth_k<{bs},*> (fused smoothing step, finest level)"
    survey_extra = (4.0 * (nnz_top - words_top) + 8.0 * bs * bs * (nnz_top - vals_top)) * dom["launches"]
    achieved_survey = (dom["alg_bytes"] + survey_extra) / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
    total_alg = sum(v["alg_bytes"] for v in prof.values())
    n_total = n_global
    value = n_total * args.steps / (ms * 1e-3)
    exchanges = int(ctx.L.uggpu_comm_exchanges(ctx.h))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(cells, top, n, args.kind, args.smoother) if world == 1 else
                   (f"3D P1 Poisson, box of {P[0]}x{P[1]}x{P[2]} unit cubes (one per GPU), Kuhn tetrahedra, base {cells * P[0]}x{cells * P[1]}x{cells * P[2]} cells, "
                    f"{top + 1} levels, {n_global} fine unknowns ({n} on rank 0), V(2,2) Jacobi damp 0.6, base solver ls+lu"),
                   "parallelism": f"dp{world}: element partition into {P[0]}x{P[1]}x{P[2]} boxes, owner-computes rows + NCCL halo copies, "
                                  f"levels <= {args.replicate_below} rows replicated" if world > 1 else "dp1",
                   "halo_exchanges_total": exchanges,
                   "cache": "inputs larger than L2 (every sweep over the finest level streams several GB; each of its vectors alone is 1 GB)",
                   "schedule": "fused", "device_bytes": dev_bytes, "setup_s": round(setup_s, 2), "preprocess_s": round(preprocess_s, 3), **({"galerkin_ms_top_down": galerkin_ms} if galerkin_ms is not None else {}),
                   "defect": [first, hist[-1]] if hist else None},
        "roofline": {"bound": "hbm", "kernel": smooth_kernel if args.smoother == "jac" else
                     f"k_dmatmul_k<{bs},2> (defect update of the smoothing step, finest level)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "launches": dom["launches"], "avg_ms": dom["ms"] / max(dom["launches"], 1),
                     "alg_bytes_per_launch": dom["alg_bytes"] / max(dom["launches"], 1),
                     "bytes_model": "as stored: 8 B per value fetched (slices of identical rows share value tables), compressed column words (DESIGN.md 2-3), vectors once",
                     "value_entries_per_entry": vals_top / max(nnz_top, 1), "stencil_slices_frac": sten_top / max((ctx.level_n(top) + 31) // 32, 1),
                     "achieved_survey_model": achieved_survey, "column_words_per_entry": words_top / max(nnz_top, 1),
                     "share_of_step": dom["ms"] / ms,
                     "cycle_alg_GBps": total_alg / (ms * 1e-3) / 1e9, "cycle_frac": total_alg / (ms * 1e-3) / 1e9 / peak},
        "kernels": prof,
        "trisolve_finest": {"launches": tri["launches"], "avg_ms": tri["ms"] / max(tri["launches"], 1),
                            "GBps": tri["alg_bytes"] / (tri["ms"] * 1e-3) / 1e9 if tri["ms"] > 0 else 0.0} if args.smoother != "jac" else None,
        "e2e": {"value": n_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": (16 * n + 8) * world,
                "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "spmv": spmv,
    }
    if world == 1 and not args.no_cpu:
        r = run_reference_cpu(args.cpu_refine, 5, args.cpu_replicas)
        kind = "reference"
        if r is None:
            r = run_port_cpu(1, 5, 5)
            kind = "port"
        line["cpu_baseline"] = {"value": r["vcycle_unknowns_per_s"], "unit": UNIT, "cores": r["cores"], "kind": kind,
                                "sample": f"{r['cycles']} V(2,2) cycles, {r['unknowns']} fine unknowns, "
                                          + (f"unmodified UG 3.12.1 numprocs (oracle/_ref/ugoracle3), {r['cores']} concurrent single-threaded "
                                             f"replicas (UG's only parallel mode is MPI), throughputs added" if kind == "reference" else "oracle/ugport.c")}
    print(json.dumps(line))
    ctx.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4, help="base cells per direction")
    ap.add_argument("--top", type=int, default=7, help="number of uniform refinements (levels - 1)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-refine", type=int, default=6, help="refinements of the host-side reference run (6 -> 274 625 unknowns)")
    ap.add_argument("--cpu-replicas", type=int, default=0, help="concurrent single-threaded reference replicas (0 = one per host core, memory permitting)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kind", default="p1", choices=["p1", "q1", "elasticity"],
                    help="p1: BASELINE configs[1] (default); q1 / elasticity: Q1 cubes, scalar / 3x3 blocks (configs[3]; use --top 6)")
    ap.add_argument("--smoother", default="jac", choices=["jac", "gs", "sgs", "sor", "ilu"],
                    help="smoother class of the cycle: jac = BASELINE configs (default); gs / sgs / sor / ilu: Gauss-Seidel family and ILU (SURVEY.md 8f.2, one GPU)")
    ap.add_argument("--galerkin", action="store_true",
                    help="replace the coarse-level matrices by the Galerkin products P^T A P (uggpu_galerkin, SURVEY.md 8f.3) before the cycle "
                         "and report the time per level in config.galerkin_ms_top_down")
    ap.add_argument("--replicate-below", type=int, default=300000,
                    help="multi-GPU: levels with at most this many rows are held completely by every rank (coarse-level gather)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
