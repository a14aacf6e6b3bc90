#!/usr/bin/env python
"""bench.py -- V(2,2)-cycle throughput of the gpuls hot path on B200 (BASELINE.json metric).

One "step" = one iteration of the linear solver `ls` with `lmgc` V(2,2) damped-Jacobi as Iter, i.e. exactly the body
of LinearSolver's loop (np/procs/ls.cc:693-708): c = 0; c = Lmgc(b) (b updated to the new defect); x += c; ||b||_2.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference [...]                        the unmodified reference on the host CPU (oracle/_ref)

Workloads:
  N = 1   C2: 3D P1 Poisson, unit cube, base 4x4x4 cells, 8 levels, 513^3 = 135 005 697 fine unknowns.
  N >= 2  C3 verbatim, a FIXED global problem ("scaling": "strong"): base 8x8x6 cells, 8 levels, 1025 x 1025 x 769 =
          807 930 625 fine unknowns, partitioned into boxes of base cells over 2x1x1 / 2x2x1 / 2x2x2 GPUs (UG's RCB element
          partition on a structured grid).  `--weak` gives every GPU a C2-sized box instead; the weak figure is also
          reported as the extra key `weak` of the default run.
Prints ONE JSON line (rank 0).  `value` = fine unknowns * K / device time of K steps with everything resident in HBM;
`e2e` = the same through the C-ABI with HOST vectors (x, b uploaded and downloaded inside the timed region, as
NP_LINEAR_SOLVER::Solver does with UG's VECTOR lists); `roofline` = the dominant kernel (fused smoothing step on the
finest level) timed with CUDA events around each of its launches inside the timed region.  Further keys: the same cycle on the
GENERAL storage path (no shared value tables, no stencil kernels) and on a VARYING-coefficient operator, Q1 Poisson and 3x3-block
elasticity, the Galerkin product, the Krylov solvers, a multi-GPU parity check against one GPU (`mgpu_parity`, N >= 2).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vcycle_unknowns_per_s"
UNIT = "unknowns/s"
ARRAYS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
WHAT = {"p1": "3D P1 Poisson, Kuhn tetrahedra (15-point rows)", "q1": "3D Q1 Poisson, hexahedra (27-point rows)",
        "elasticity": "3D Q1 linear elasticity (3x3 blocks, 27 block entries per row), hexahedra",
        "p1var": "3D P1 diffusion with a smoothly VARYING coefficient (no two rows share values), Kuhn tetrahedra"}
SMOOTHER_TXT = {"jac": "Jacobi damp 0.6", "gs": "Gauss-Seidel damp 1.0", "sgs": "symmetric Gauss-Seidel damp 1.0", "sor": "SOR omega 1.1",
                "ilu": "ILU(0) beta 0 damp 1.0"}


def workload_name(kind, cells, top, n_global, smoother="jac", P=(1, 1, 1)):
    nn = [c * 2 ** top + 1 for c in cells]
    part = "" if P == (1, 1, 1) else f", partitioned into {P[0]}x{P[1]}x{P[2]} boxes of base cells (one per GPU)"
    note = " (the Kuhn hierarchy has 15 connections per row; UG's own tetrahedron rule gives 14.6 on average, DESIGN.md 4)" if kind in ("p1", "p1var") else ""
    return (f"{WHAT[kind]}, box of {cells[0]}x{cells[1]}x{cells[2]} base cells, {top + 1} levels, {nn[0]}x{nn[1]}x{nn[2]} nodes = "
            f"{n_global} fine unknowns{part}, V(2,2) {'block-' if kind == 'elasticity' else ''}{SMOOTHER_TXT[smoother]}, base solver ls+lu{note}")


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


def equal_size_inside_ug(refine: int, cycles: int, exe_name: str = "ugoracle3", grid_args=None, bs: int = 1):
    """The gpuls numprocs INSIDE the unmodified UG (oracle/_ref/ugoracle3 --gpu: PreProcess flattens the VECTOR/MATRIX lists, Solver
    uploads x and b, runs the cycles on the device, scatters x, b, c back into the VVALUEs) next to UG's own CPU numprocs on the SAME
    hierarchy in the same process: wall time of NP_LINEAR_SOLVER::Solver on both sides."""
    exe = os.path.join(ROOT, "oracle", "_ref", exe_name)
    lib = os.environ.get("UGGPU_BENCH_LIB") or os.path.join(ROOT, "ug_b200", "lib", "libuggpu.so")      # (override: the CPU stand-in, to test this parser without a GPU)
    if not os.path.exists(exe):
        return None
    out = subprocess.run([exe] + (grid_args or ["--grid", "tet", "--refine", str(refine), "--damp", "0.6"]) + ["--cycles", str(cycles), "--gpu", lib, "--nokrylov"],
                         capture_output=True, text=True, timeout=900)
    n = None
    m = re.search(r"n=\[([0-9,]+)\]", out.stdout)
    if m:
        n = int(m.group(1).split(",")[-1]) * bs
    for line in out.stdout.splitlines():
        if line.startswith(("PASS", "FAIL")) and "device base solver" in line:
            mg, mc = re.search(r"t_gpu=([0-9.eE+-]+)s", line), re.search(r"t_cpu=([0-9.eE+-]+)s", line)
            me = re.search(r"relerr x=([0-9.eE+-]+)", line)
            if mg and mc and n:
                tg, tc = float(mg.group(1)), float(mc.group(1))
                pg, pc = re.search(r"pre_gpu=([0-9.eE+-]+)s", line), re.search(r"pre_cpu=([0-9.eE+-]+)s", line)
                return {"unknowns": n, "cycles": cycles, "gpu_numprocs_inside_ug_s": tg, "cpu_numprocs_s": tc, "ratio": tc / tg if tg > 0 else None,
                        "preprocess_gpu_s": float(pg.group(1)) if pg else None, "preprocess_cpu_s": float(pc.group(1)) if pc else None,
                        "gpu_unknowns_per_s": n * cycles / tg if tg > 0 else None, "cpu_unknowns_per_s": n * cycles / tc if tc > 0 else None,
                        "parity": line.split(" ", 1)[0], "relerr_x": float(me.group(1)) if me else None,
                        "what": "wall time of NP_LINEAR_SOLVER::Solver: gpuls+gpulmgc+gpujac+gputransfer (device base solver; x, b up, x, b, c down through "
                                "the VECTOR lists inside the call) vs ls+lmgc+jac+transfer, same UG hierarchy, same process, 1 host core"}
    return {"error": (out.stdout[-300:] + out.stderr[-300:])}


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
    nn = round(n ** (1 / 3))
    sample = (f"{'UG 3.12.1 ls+lmgc+jac+transfer' if kind == 'reference' else 'oracle/ugport.c'}: {r['cycles']} V(2,2) cycles on a "
              f"{r.get('levels', '?')}-level unit-cube tet hierarchy with {n} fine unknowns (largest the host builds in ~10 s), "
              f"{r['cores']} concurrent single-threaded replica(s), throughputs added")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["vcycle_unknowns_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["cycles"], "warmup": 0, "ms_per_step": r["s_per_cycle"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"3D P1 Poisson, unit cube, tetrahedra (UG's own refinement), {r.get('levels', '?')} levels, {nn}x{nn}x{nn} nodes = {n} fine unknowns "
                                f"per replica x {r['cores']} replicas, V(2,2) Jacobi damp 0.6, base solver ls+lu -- the SIZE THE HOST CAN BUILD, not the GPU arm's "
                                f"problem (UG's grid manager needs ~2.5 kB per unknown: 513^3 would take ~350 GB)"),
                   "same_size_as_gpu_arm": False, "measured_on": sample},
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


KINDS = {"smooth": 0, "jac": 1, "restrict": 2, "interpolate": 3, "vecop": 4, "reduce": 5, "dmatmul": 6, "base": 7, "trisolve": 8, "halo": 9, "allreduce": 10}


class Env:
    """Environment switches of libuggpu.so (read with getenv at set-up / launch time) for the duration of one workload."""

    def __init__(self, env):
        self.env, self.old = env or {}, {}

    def __enter__(self):
        for k, v in self.env.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def nonzero_w(sten_w, kind):
    """Nonzero coefficients of the dominant stencil of the synthetic operators (stx.cu sten_nonzero): the P1 Laplacian on a Kuhn mesh has 8
    zero coefficients among its 15 connections (the diagonal neighbours), the Q1 Laplacian 6 among 27 (the face neighbours)."""
    if os.environ.get("UGGPU_KEEP_ZERO_ENTRIES"):
        return sten_w
    return {("p1", 15): 7, ("q1", 27): 21}.get((kind, sten_w), sten_w)


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure(spec, rt):
    """Builds the hierarchy of `spec` on all ranks, times spec['steps'] solver iterations, returns the figures (same dict on every
    rank as far as rank 0 needs them).  rt = (rank, world, local, torch, dist)."""
    import numpy as np
    from ug_b200 import capi, mgpu
    rank, world, local, torch, dist = rt
    kind, cells, top, P = spec["kind"], spec["cells"], spec["top"], spec["P"]
    steps, warmup = spec["steps"], spec["warmup"]
    smoother = spec.get("smoother", "jac")
    peak, peak_src = peak_hbm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with Env(spec.get("env")):
        ctx = capi.Context(local)
        if world > 1:
            mgpu.init_comm(ctx, rank, world)
        A = ctx.handle("A")
        t0 = time.perf_counter()
        ctx.call("uggpu_synth_hierarchy_part", mgpu.KINDS[kind], cells[0], cells[1], cells[2], top, A, P[0], P[1], P[2], rank, C.c_int64(spec.get("replicate_below", 300000)))
        bs = ctx.level_bs(top)
        n = ctx.level_n(top) * bs                      # unknowns (UG counts vector components), this rank
        n_global = int(ctx.L.uggpu_level_n_global(ctx.h, top)) * bs
        for name in ("x", "b", "c"):
            for l in range(top + 1):
                ctx.alloc(l, name)
        ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
        sm_damp = {"jac": 0.6, "gs": 1.0, "sgs": 1.0, "sor": 1.1, "ilu": 1.0}[smoother]
        cfg = ctx.lmgc_cfg(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=sm_damp, fused=1, smoother=smoother)
        out = {"kind": kind, "n_global": n_global, "n_rank0": n, "bs": bs}
        if spec.get("galerkin"):      # setup operation, outside the timed steps: A_{l-1} := P^T A_l P cascaded from the top level down (uggpu_galerkin)
            gms, gk = [], []
            for l in range(top, 0, -1):
                ctx.sync()
                ctx.call("uggpu_prof_enable", 1)
                tg = time.perf_counter()
                ctx.call("uggpu_galerkin", l, A)
                ctx.sync()
                gms.append(round((time.perf_counter() - tg) * 1e3, 3))
                nl, kms, kb = C.c_int64(0), C.c_double(0), C.c_double(0)
                ctx.call("uggpu_prof_summary", 12, l, C.byref(nl), C.byref(kms), C.byref(kb))      # UGGPU_K_GALERKIN: the product kernel alone (CUDA events)
                ctx.call("uggpu_prof_enable", 0)
                gk.append(round(kms.value, 3))
            out["galerkin_ms_top_down"] = gms             # whole calls (host clock): sort of the transposed stencil, allocations, product, tables of the new matrix
            out["galerkin_kernel_ms_top_down"] = gk
            nf, zf = ctx.level_n(top), int(ctx.L.uggpu_mat_nnz(ctx.h, top, A))
            nc, zc = ctx.level_n(top - 1), int(ctx.L.uggpu_mat_nnz(ctx.h, top - 1, A))
            zp = int(ctx.L.uggpu_transfer_nnz(ctx.h, top, 0))
            # compulsory bytes of the finest product: fine matrix once (12 B per entry), P twice (gathered rows), coarse matrix written
            gbytes = 12.0 * zf * bs * bs + 2 * 12.0 * zp + 12.0 * zc * bs * bs + 4.0 * (nf + nc)
            gt = gk[0] if gk[0] > 0 else gms[0]
            out["galerkin_finest"] = {"kernel_ms": gk[0], "call_ms": gms[0], "alg_bytes": gbytes, "GBps": gbytes / (gt * 1e-3) / 1e9, "frac": gbytes / (gt * 1e-3) / 1e9 / peak}
        ctx.sync()
        t_pre = time.perf_counter()
        ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
        ctx.sync()
        out["preprocess_s"] = round(time.perf_counter() - t_pre, 3)      # LmgcPreProcess: base-level LU, Gauss-Seidel schedules / ILU decompositions
        out["setup_s"] = round(time.perf_counter() - t0, 2)
        X, B, Cc = ctx.handle("x"), ctx.handle("b"), ctx.handle("c")
        res = capi.LResult()
        ctx.call("uggpu_ls_defect", 0, top, X, B, A)
        ctx.call("uggpu_ls_residuum", 0, top, B, C.byref(res))
        first = res.last_defect[0]
        absl, red = capi._vs([1e-300]), capi._vs([1e-300])

        hbuf = np.zeros(max(steps, warmup, spec.get("e2e_solve_cycles", 10)) * bs)

        def step(k=1):
            # k iterations of LinearSolver's loop (ls.cc:693-708) in ONE call, as NP_LINEAR_SOLVER::Solver runs them: the defect norm comes back
            # to the host after every iteration (convergence test), nothing else does
            ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, X, B, A, Cc, k, absl, red, C.byref(res), hbuf.ctypes.data_as(C.POINTER(C.c_double)))

        stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
        step(warmup)
        hist = []
        barrier()
        sampler = ClockSampler(local) if (rank == 0 and spec.get("clocks")) else None
        ctx.call("uggpu_prof_enable", 1)
        launches0 = ctx.launch_count()
        exch0 = int(ctx.L.uggpu_comm_exchanges(ctx.h))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step(steps)
        hist = [float(hbuf[i * bs]) for i in range(steps)]
        e1.record(stream)
        barrier()
        ms = allmax(e0.elapsed_time(e1))
        ms_events = ms
        out["launches"] = ctx.launch_count() - launches0
        out["halo_exchanges_per_step"] = (int(ctx.L.uggpu_comm_exchanges(ctx.h)) - exch0) / steps
        prof = {}
        cnt, kms, kby = C.c_int64(), C.c_double(), C.c_double()
        for name, k in KINDS.items():
            ctx.call("uggpu_prof_summary", k, -1, C.byref(cnt), C.byref(kms), C.byref(kby))
            prof[name] = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
        domk = 0 if smoother == "jac" else 6      # the fused smoothing step (Jacobi) / the defect update dmatmul_minus (Gauss-Seidel family), finest level
        ctx.call("uggpu_prof_summary", domk, top, C.byref(cnt), C.byref(kms), C.byref(kby))
        dom = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
        ctx.call("uggpu_prof_summary", 8, top, C.byref(cnt), C.byref(kms), C.byref(kby))
        tri = {"launches": cnt.value, "ms": kms.value, "alg_bytes": kby.value}
        ctx.call("uggpu_prof_enable", 0)
        # the K timed steps proper: without the per-kernel events (two cudaEventRecord around every launch are what a user does not have)
        barrier()
        e0.record(stream)
        step(steps)
        e1.record(stream)
        barrier()
        ms = allmax(e0.elapsed_time(e1))
        out["ms_per_step_with_kernel_events"] = ms_events / steps
        out["ms"] = ms
        out["ms_per_step"] = ms / steps
        out["value"] = n_global * steps / (ms * 1e-3)
        out["kernels"] = prof
        out["kernel_sum_ms_per_step"] = sum(v["ms"] for v in prof.values()) / steps
        if world > 1:      # every rank's view (rank 0's is `kernels`): a rank that waits for a slower neighbour shows it in halo / allreduce / the comm kernels
            mine = {k: round(v["ms"] / steps, 3) for k, v in prof.items() if v["ms"] > 0}
            mine["dominant_avg_ms"] = round(dom["ms"] / max(dom["launches"], 1), 4)
            mine["rows"] = ctx.level_n(top)
            allr = [None] * world
            dist.all_gather_object(allr, mine)
            out["per_rank_kernels_ms_per_step"] = allr
        out["defect"] = [first, hist[-1]] if hist else None
        out["transport"] = ctx.halo_transport() if world > 1 else "none (one GPU)"
        out["device_bytes"] = ctx.device_bytes()

        # roofline of the dominant kernel -----------------------------------------------------------------------------
        nrows_top = ctx.level_n(top)
        nnz_top, words_top = int(ctx.L.uggpu_mat_nnz(ctx.h, top, A)), int(ctx.L.uggpu_mat_col_words(ctx.h, top, A))
        vals_top = int(ctx.L.uggpu_mat_val_entries(ctx.h, top, A))      # entries whose values a sweep fetches (shared value tables)
        sten_top = int(ctx.L.uggpu_mat_stencil_slices(ctx.h, top, A)) if not os.environ.get("UGGPU_NO_STENCIL") else 0
        sten_w = round(nnz_top / max(nrows_top, 1))
        stx = not os.environ.get("UGGPU_NO_STX") and nrows_top >= (1 << 20)
        if sten_top > 0 and bs == 1 and sten_w in (15, 27):
            smooth_kernel = (f"k_smooth_stx<*,W> + k_smooth_xrows (fused smoothing step: rows that are exactly the {sten_w}-entry stencil -- the kernel runs on its nonzero coefficients, W = {nonzero_w(sten_w, kind)} -- + the packed exception rows, two kernels side by side, finest level)"
                             if stx else f"k_smooth_sten<*,{sten_w}> (fused smoothing step, stencil variant, finest level)")
        elif sten_top > 0 and bs == 3:
            smooth_kernel = ("k_smooth_stx3<*> + k_smooth_xrows<3,*> (fused smoothing step, 3x3 blocks: stencil rows + packed exception rows, finest level)"
                             if stx else "k_smooth_sten3<*> (fused smoothing step, 3x3-block stencil variant, finest level)")
        else:
            smooth_kernel = f"k_smooth_k<{bs},*> (fused smoothing step, finest level)"
        achieved = dom["alg_bytes"] / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
        # SURVEY.md 8(d) counts (8 b^2 + 4) bytes per entry + 4 (n + 1); the stored form fetches what uggpu_mat_pass_bytes says
        pass_bytes = float(ctx.L.uggpu_mat_pass_bytes(ctx.h, top, A))
        survey_extra = ((8.0 * bs * bs + 4.0) * nnz_top + 4.0 * (nrows_top + 1) - pass_bytes) * dom["launches"]
        achieved_survey = (dom["alg_bytes"] + survey_extra) / (dom["ms"] * 1e-3) / 1e9 if dom["ms"] > 0 else 0.0
        total_alg = sum(v["alg_bytes"] for v in prof.values())
        out["roofline"] = {"bound": "hbm", "kernel": smooth_kernel if smoother == "jac" else f"k_dmatmul_k<{bs},2> (defect update of the smoothing step, finest level)",
                           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                           "launches": dom["launches"], "avg_ms": dom["ms"] / max(dom["launches"], 1),
                           "alg_bytes_per_launch": dom["alg_bytes"] / max(dom["launches"], 1),
                           "bytes_model": "as stored: the matrix bytes a pass fetches (uggpu_mat_pass_bytes: row mask + packed exception rows for stencil matrices; values + compressed column words otherwise, DESIGN.md 2-3) + every vector once",
                           "matrix_bytes_per_pass": pass_bytes,
                           "value_entries_per_entry": vals_top / max(nnz_top, 1), "stencil_slices_frac": sten_top / max((nrows_top + 31) // 32, 1),
                           "achieved_survey_model": achieved_survey, "frac_survey_model": achieved_survey / peak,
                           "survey_model": "SURVEY.md 8(d): 8 b^2 + 4 B per matrix entry whatever the storage; above 1 means the matrix stream no longer crosses HBM",
                           "column_words_per_entry": words_top / max(nnz_top, 1), "share_of_step": dom["ms"] / ms_events,
                           "cycle_alg_GBps": total_alg / (ms * 1e-3) / 1e9, "cycle_frac": total_alg / (ms * 1e-3) / 1e9 / peak}
        out["trisolve_finest"] = ({"launches": tri["launches"], "avg_ms": tri["ms"] / max(tri["launches"], 1),
                                   "GBps": tri["alg_bytes"] / (tri["ms"] * 1e-3) / 1e9 if tri["ms"] > 0 else 0.0} if smoother != "jac" else None)

        # ---- end to end: host vectors in, host vectors out, every step -------------------------------------------------
        if spec.get("e2e_steps", 0) > 0:
            xh = torch.zeros(n, dtype=torch.float64).pin_memory()      # n counts components: the vectors hold n doubles
            bh = torch.empty(n, dtype=torch.float64).pin_memory()
            ctx.call("uggpu_vec_download", top, X, C.c_void_p(xh.data_ptr()))
            ctx.call("uggpu_vec_download", top, B, C.c_void_p(bh.data_ptr()))

            def e2e_step(k=1):
                # b first (the cycle starts with it); x follows on the copy stream and hides behind the cycle, whose last kernel
                # is the first to touch it
                ctx.call("uggpu_vec_upload", top, B, C.c_void_p(bh.data_ptr()))
                ctx.call("uggpu_vec_upload_async", top, X, C.c_void_p(xh.data_ptr()))
                ctx.call("uggpu_ls_residuum", 0, top, B, C.byref(res))
                step(k)
                ctx.call("uggpu_vec_download", top, X, C.c_void_p(xh.data_ptr()))
                ctx.call("uggpu_vec_download", top, B, C.c_void_p(bh.data_ptr()))

            e2e_step()
            barrier()
            t1 = time.perf_counter()
            for _ in range(spec["e2e_steps"]):
                e2e_step()
            barrier()
            e2e_s = allmax((time.perf_counter() - t1) / spec["e2e_steps"])
            out["e2e"] = {"value": n_global / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 16 * n * world, "d2h_bytes_per_step": (16 * n + 8) * world,
                          "ms_per_step": e2e_s * 1e3, "steps": spec["e2e_steps"]}
            # what NP_LINEAR_SOLVER::Solver does: x, b up once, k cycles on the device, x, b down once
            kk = spec.get("e2e_solve_cycles", 10)
            barrier()
            t1 = time.perf_counter()
            e2e_step(kk)
            barrier()
            es = allmax(time.perf_counter() - t1)
            out["e2e_solve"] = {"value": n_global * kk / es, "unit": UNIT, "cycles": kk, "ms_per_cycle": es * 1e3 / kk, "ms_total": es * 1e3,
                                "h2d_bytes": 16 * n * world, "d2h_bytes": (16 * n + 8) * world,
                                "what": "one solve as the gpuls numproc runs it: x, b host->device once, k cycles resident, x, b device->host once"}
        if sampler:
            out["clocks"] = sampler.stop()

        # BASELINE.json's second figure, "SpMV HBM GB/s vs peak": the plain dmatmul_minus (c -= A x) on the finest level, timed by the
        # library's per-kernel events after everything else (c is work space: every cycle starts by overwriting it)
        if spec.get("spmv"):
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
                    gb = b2.value / (m2.value * 1e-3) / 1e9
                    extra = ((8.0 * bs * bs + 4.0) * nnz_top + 4.0 * (nrows_top + 1) - pass_bytes) * c2.value
                    out["spmv"] = {"kernel": ((f"k_dmatmul_stx<2,W = {nonzero_w(sten_w, kind)} of {sten_w}> + k_dmatmul_xrows" if stx else f"k_dmatmul_sten<2,{sten_w}>") if (sten_top > 0 and bs == 1 and sten_w in (15, 27)) else f"k_dmatmul_k<{bs},2>") + " (x -= A y, finest level)",
                                   "launches": int(c2.value), "avg_ms": m2.value / c2.value, "alg_bytes_per_launch": b2.value / c2.value, "GBps": gb, "frac": gb / peak,
                                   "GBps_survey_model": (b2.value + extra) / (m2.value * 1e-3) / 1e9, "frac_survey_model": (b2.value + extra) / (m2.value * 1e-3) / 1e9 / peak}
            except Exception as e:          # a reported figure, never a reason to lose the bench line
                out["spmv"] = {"error": str(e)[:200]}

        # Krylov accelerators around the cycle: time per iteration, resident
        if spec.get("krylov"):
            try:
                kr = {}
                for l in range(top + 1):
                    for name in ("p", "t2", "r", "v", "s", "q"):
                        ctx.alloc(l, name)
                for solver in ("cg", "bcgs"):
                    ctx.call("uggpu_dset", 0, top, 0, X, C.c_double(0.0))
                    ctx.call("uggpu_synth_rhs", top, B)
                    ctx.call("uggpu_ls_residuum", 0, top, B, C.byref(res))
                    d0 = res.last_defect[0]
                    its = 4
                    barrier()
                    t1 = time.perf_counter()
                    if solver == "cg":
                        ctx.call("uggpu_cg_solve", C.byref(cfg), 0, top, X, B, A, Cc, ctx.handle("p"), ctx.handle("t2"), its, absl, red, C.byref(res), None)
                    else:
                        work = (C.c_int * 6)(*[ctx.handle(nm) for nm in ("r", "p", "v", "s", "t2", "q")])
                        ctx.call("uggpu_bcgs_solve", C.byref(cfg), 0, top, X, B, A, work, capi._vs([1.0]), 0, its, absl, red, C.byref(res), None)
                    barrier()
                    dt = allmax(time.perf_counter() - t1)
                    nit = max(int(res.number_of_linear_iterations), 1)
                    kr[solver] = {"iterations": nit, "ms_per_iteration": dt * 1e3 / nit, "cycles_per_iteration": 1, "defect": [d0, res.last_defect[0]],
                                  "unknowns_per_s": n_global * nit / dt}
                out["krylov"] = kr
            except Exception as e:
                out["krylov"] = {"error": str(e)[:200]}
        ctx.close()
    return out


def kuhn_elements(nx, ny, nz):
    """Rows of the corner vectors of the 6 Kuhn tetrahedra (diagonal 0-6, UG's hexahedron corner numbering) of every cube of a structured
    nx*ny*nz-cell grid with lexicographic rows (x fastest) -- the mesh uggpu_synth_hierarchy assembles analytically -- in cube order."""
    import numpy as np
    NX, NY = nx + 1, ny + 1
    i, j, k = np.meshgrid(np.arange(nx, dtype=np.int64), np.arange(ny, dtype=np.int64), np.arange(nz, dtype=np.int64), indexing="ij")
    base = (i + NX * (j + NY * k)).transpose(2, 1, 0).ravel()
    off = np.array([0, 1, 1 + NX, NX, NX * NY, 1 + NX * NY, 1 + NX + NX * NY, NX + NX * NY], dtype=np.int64)
    tets = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]])
    er = (base[:, None, None] + off[tets][None, :, :]).astype(np.int32).reshape(-1)
    return np.arange(0, er.size + 1, 4, dtype=np.int64), er


def assemble_bench(local, cells, top):
    """SURVEY.md 8f.4 at size: uggpu_assemble (the element loop of np/procs/assemble.cc:657 + AssembleDirichletBoundary) on the finest level
    of the synthetic P1 hierarchy from its element list, into a second matrix with the same pattern; kernel time from CUDA events
    (uggpu_prof), compared entry by entry with the matrix the generator wrote analytically.  A setup operation, outside the timed steps."""
    import numpy as np
    from ug_b200 import capi, mgpu
    peak, _ = peak_hbm()
    ctx = capi.Context(local)
    try:
        A = ctx.handle("A")
        ctx.call("uggpu_synth_hierarchy", mgpu.KINDS["p1"], cells[0], cells[1], cells[2], top, A)
        n = ctx.level_n(top)
        nn = [c * 2 ** top + 1 for c in cells]
        assert n == nn[0] * nn[1] * nn[2]
        t0 = time.perf_counter()
        ep, er = kuhn_elements(nn[0] - 1, nn[1] - 1, nn[2] - 1)
        h = 1.0 / (nn[0] - 1)
        z, y, x = np.meshgrid(np.arange(nn[2]), np.arange(nn[1]), np.arange(nn[0]), indexing="ij")
        coord = np.stack([x.ravel() * h, y.ravel() * h, z.ravel() * h], axis=1).astype(np.float64)
        bnd = ((x == 0) | (x == nn[0] - 1) | (y == 0) | (y == nn[1] - 1) | (z == 0) | (z == nn[2] - 1)).ravel()
        skip = bnd.astype(np.uint32)
        del x, y, z
        host_s = time.perf_counter() - t0
        nnz = int(ctx.L.uggpu_mat_nnz(ctx.h, top, A))
        rowptr = np.zeros(n + 1, np.int32); col = np.zeros(nnz, np.int32); val = np.zeros(nnz)
        ctx.call("uggpu_mat_get", top, A, capi._p(rowptr), capi._p(col), capi._p(val))
        K = ctx.handle("K")
        ctx.call("uggpu_mat_set_pattern", top, K, capi._p(rowptr), capi._p(col))
        ctx.alloc(top, "x"); ctx.alloc(top, "b"); ctx.alloc(top, "b0")
        ctx.call("uggpu_dset", top, top, 0, ctx.handle("x"), 0.0)
        ctx.call("uggpu_synth_rhs", top, ctx.handle("b0"))
        fe = dict(problem=0, dim=3, E=1.0, nu=0.3, source=[1.0])
        ms = []
        for rep in range(3):
            ctx.call("uggpu_prof_enable", 1)
            t1 = time.perf_counter()
            ctx.assemble(top, "x", "b", "K", fe, ep, er, None, coord, skip)
            ctx.sync()
            call_ms = (time.perf_counter() - t1) * 1e3
            launches, kms, byt = C.c_int64(0), C.c_double(0), C.c_double(0)
            ctx.call("uggpu_prof_summary", 11, top, C.byref(launches), C.byref(kms), C.byref(byt))
            ctx.call("uggpu_prof_enable", 0)
            ms.append((kms.value, call_ms))
        kval = ctx.mat_values(top, "K", nnz)
        b, b0 = ctx.get(top, "b"), ctx.get(top, "b0")
        scale = float(np.abs(val).max())
        nelem = ep.size - 1
        # compulsory bytes: the element list once (16 B per tetrahedron) + 4 corner coordinates per element (96 B, gathered) + the matrix
        # written once (12 B per entry as stored explicitly) + rhs, x, skip
        alg = 16.0 * nelem + 96.0 * nelem + 8.0 * nnz + 4.0 * nnz + 20.0 * n
        best = min(m[0] for m in ms)
        return {"workload": f"P1 Poisson, {nn[0]}x{nn[1]}x{nn[2]} nodes = {n} rows, {nelem} Kuhn tetrahedra, {nnz} matrix entries",
                "kernel_ms": round(best, 3), "call_ms_with_uploads_and_sort": round(min(m[1] for m in ms), 1), "host_element_list_s": round(host_s, 2),
                "elements_per_s": nelem / (best * 1e-3), "alg_bytes": alg, "GBps": alg / (best * 1e-3) / 1e9, "frac": alg / (best * 1e-3) / 1e9 / peak,
                "max_abs_diff_vs_generator_matrix_rel": float(np.abs(kval - val).max() / scale), "entries_bit_identical_frac": float(np.mean(kval == val)),
                "rhs_max_abs_diff_rel": float(np.abs(b - b0).max() / np.abs(b0).max()),
                "parity": "bit-exact against the reference's NP_LOCAL_ASSEMBLE loop on 6 hierarchies (tests/test_gpu_parity.py::test_gpu_assemble_bitexact, "
                          "tests/test_dropin.py assemble-*); here: agreement with the generator's analytic stencil"}
    finally:
        ctx.close()



def brief(m, keys=("value", "ms_per_step", "n_global", "launches", "kernel_sum_ms_per_step", "defect", "device_bytes", "setup_s", "transport", "halo_exchanges_per_step")):
    """What an extra workload contributes to the JSON line."""
    r = {k: m[k] for k in keys if k in m}
    rf = m["roofline"]
    r["dominant_kernel"] = {k: rf[k] for k in ("kernel", "avg_ms", "achieved", "frac", "achieved_survey_model", "frac_survey_model", "alg_bytes_per_launch", "value_entries_per_entry",
                                               "column_words_per_entry", "cycle_frac")}
    r["kernels_ms_per_step"] = {k: round(v["ms"] / max(m.get("steps", 1), 1), 4) for k, v in m["kernels"].items() if v["ms"] > 0}
    for k in ("galerkin_ms_top_down", "galerkin_kernel_ms_top_down", "galerkin_finest", "spmv", "krylov"):
        if k in m:
            r[k] = m[k]
    return r


def our_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gpuls path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = ARRAYS.get(world)
    if P is None:
        raise SystemExit(f"bench.py: unsupported GPU count {world} (1, 2, 4 or 8)")
    rt = (rank, world, local, torch, dist)
    from ug_b200 import mgpu

    # ---- multi-GPU parity, before anything is timed: the partitioned solve must be the one-GPU solve bit for bit ------------------
    parity = None
    if world > 1 and not args.no_parity:
        parity = [mgpu.parity_check(rank, world, local, kind=k, top=t, fused=f, small_levels=sm)
                  for k, t, f, sm in (("p1", 5, 1, True), ("p1", 5, 1, False), ("p1", 4, 0, False), ("q1", 4, 1, True), ("elasticity", 4, 1, True))]
        if not all(p["ok"] for p in parity):
            if rank == 0:
                print(json.dumps({"metric": METRIC, "error": "multi-GPU parity check failed", "mgpu_parity": parity}))
            return 3

    # ---- the headline workload ----------------------------------------------------------------------------------------------------
    strong = world > 1 and not args.weak and args.kind == "p1" and args.cells == 4 and args.top == 7
    if strong:
        cells = (8, 8, 6)                                   # C3 verbatim: a fixed global problem over 2 / 4 / 8 GPUs
    elif world > 1:
        cells = (args.cells * P[0], args.cells * P[1], args.cells * P[2])
    else:
        cells = (args.cells,) * 3
    steps = args.steps
    main = dict(kind=args.kind, cells=cells, top=args.top, P=P, steps=steps, warmup=args.warmup, smoother=args.smoother, e2e_steps=args.e2e_steps,
                clocks=True, spmv=True, galerkin=args.galerkin, replicate_below=args.replicate_below, krylov=args.krylov)
    m = measure(main, rt)
    m["steps"] = steps
    extras = {}
    xs, xw = max(3, min(5, steps)), 3

    def extra(name, **kw):
        spec = dict(kind="p1", cells=cells, top=args.top, P=P, steps=xs, warmup=xw, replicate_below=args.replicate_below)
        spec.update(kw)
        try:
            r = measure(spec, rt)
            r["steps"] = spec["steps"]
            extras[name] = brief(r)
            extras[name]["workload"] = workload_name(spec["kind"], spec["cells"], spec["top"], r["n_global"], spec.get("smoother", "jac"), P)
            if spec.get("env"):
                extras[name]["env"] = spec["env"]
        except Exception as e:          # an extra figure is never a reason to lose the bench line
            extras[name] = {"error": str(e)[:300]}

    if not args.no_extras and args.kind == "p1" and args.smoother == "jac":
        if world == 1:
            # the same workload on the general storage path: what an unstructured / variable-coefficient matrix gets
            extra("shared_tables_generic_kernel", env={"UGGPU_NO_STENCIL": "1"})
            extra("general_path", env={"UGGPU_NO_SHARED_VALUES": "1", "UGGPU_NO_STENCIL": "1"}, spmv=True)
            extra("varying_coefficient", kind="p1var", spmv=True)
            extra("q1_poisson", kind="q1")
            extra("elasticity_3x3", kind="elasticity", top=args.top - 1)
            extra("galerkin", galerkin=True, steps=3)
            extra("krylov", krylov=True, steps=3)
            try:
                extras["assemble"] = assemble_bench(local, (args.cells,) * 3, args.top - 1)
            except Exception as e:
                extras["assemble"] = {"error": str(e)[:300]}
        else:
            c4 = (4, 4, 4)
            extra("q1_poisson_strong", kind="q1", cells=c4)
            extra("elasticity_3x3_strong", kind="elasticity", cells=c4, top=args.top - 1)
            extra("varying_coefficient_strong", kind="p1var", cells=c4)
            if strong:
                extra("weak", cells=(4 * P[0], 4 * P[1], 4 * P[2]))

    if rank != 0:
        return 0
    n_total = m["n_global"]
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.kind, cells, args.top, n_total, args.smoother, P),
                   "parallelism": (f"dp{world}: element partition into {P[0]}x{P[1]}x{P[2]} boxes, owner-computes rows + ghost columns, halo transport: {m['transport']}, "
                                   f"levels <= {args.replicate_below} rows held completely by every rank" if world > 1 else "dp1"),
                   "scaling_note": ("strong: the global problem (SURVEY.md C3) is the same at N = 2, 4, 8; N = 1 runs C2 (135 M unknowns: C3 does not fit one GPU), so "
                                    "efficiency against N = 1 compares per-GPU throughput" if strong else
                                    ("weak: every GPU holds a C2-sized box" if world > 1 else "one GPU")),
                   "halo_exchanges_per_step": m["halo_exchanges_per_step"], "halo_transport": m["transport"],
                   "cache": "inputs larger than L2 (every sweep over the finest level streams several GB; each of its vectors alone is >= 0.8 GB)",
                   "schedule": "fused", "device_bytes": m["device_bytes"], "setup_s": m["setup_s"], "preprocess_s": m["preprocess_s"],
                   "ms_per_step_with_kernel_events": m.get("ms_per_step_with_kernel_events"), "kernel_sum_ms_per_step": m["kernel_sum_ms_per_step"],
                   **({"galerkin_ms_top_down": m["galerkin_ms_top_down"], "galerkin_finest": m["galerkin_finest"]} if "galerkin_ms_top_down" in m else {}),
                   "defect": m["defect"]},
        "roofline": m["roofline"],
        "kernels": m["kernels"],
        "trisolve_finest": m["trisolve_finest"],
        "e2e": m.get("e2e"),
        "e2e_solve": m.get("e2e_solve"),
        "per_rank_kernels_ms_per_step": m.get("per_rank_kernels_ms_per_step"),
        "gpu_launches": m["launches"],
        "clocks": m.get("clocks"),
        "spmv": m.get("spmv"),
    }
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:        # ncu capture of the default workload only
        tj = json.load(open(tpath))
        if tj.get("workload") == {"kind": args.kind, "cells": args.cells, "top": args.top}:
            line["roofline"]["traffic"] = tj.get("dominant_kernel_dram_bytes_per_launch", tj.get("k_smooth_k_dram_bytes_per_launch"))
    if "krylov" in m:
        line["krylov"] = m["krylov"]
    if parity is not None:
        line["mgpu_parity"] = {"ok": True, "x_bitexact": all(p["x_bitexact"] for p in parity), "b_bitexact": all(p["b_bitexact"] for p in parity),
                               "hist_relerr": max(p["hist_relerr"] for p in parity), "cases": parity}
    line.update(extras)
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
        if not args.no_extras:
            try:
                line["equal_size_inside_ug"] = equal_size_inside_ug(args.cpu_refine, 5)
            except Exception as e:
                line["equal_size_inside_ug"] = {"error": str(e)[:200]}
            # BASELINE.json's other configurations, as UG itself builds them, GPU numprocs inside UG next to the CPU numprocs: C1 verbatim (2D P1,
            # 6 refinements, 4 225 unknowns -- launch-bound on a GPU), a C4-type hierarchy (Q1 hexahedra, 3x3 blocks) and a C5-type one
            # (adaptively refined tetrahedra: partial levels, irregular rows) at the sizes the host builds in seconds
            for key, exe_name, ga, bsz in (("c1_inside_ug", "ugoracle2", ["--grid", "tri", "--refine", "6", "--damp", "0.8"], 1),
                                           ("c4_inside_ug", "ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "4", "--damp", "0.6"], 3),
                                           ("c5_inside_ug", "ugoracle3", ["--grid", "tet", "--refine", "4", "--adapt", "2", "--damp", "0.6"], 1),
                                           # algebraic levels: 65^3 on 33^3 (collapsed to level 0) on four levels built by the reference's
                                           # selectionAMG in every PreProcess (host, both sides); the cycle over all of them on the device vs on the host
                                           ("amg_inside_ug", "ugoracle3", ["--grid", "tet", "--refine", "5", "--collapse", "--refine2", "1", "--damp", "0.6", "--amg", "selectionAMG",
                                                                           "$strongRel 0.25 $C Greedy $I Average $CM Galerkin $vectLimit 40"], 1),
                                           # the same size with the algebraic levels built by the device library itself (gputransfer $gpuamg VanekPC: aggregation,
                                           # piecewise constant interpolation, Galerkin matrices; levels on the device only) against the reference's clusterAMG
                                           # numproc on the host: preprocess_* are the PreProcess brackets, i.e. the two AMG setups
                                           ("gpuamg_inside_ug", "ugoracle3", ["--grid", "tet", "--refine", "5", "--collapse", "--refine2", "1", "--damp", "0.6", "--amg", "clusterAMG",
                                                                              "$strongVanek 0.08 $C VanekNeuss $I PiecewiseConstant $CM Galerkin $vectLimit 60",
                                                                              "--gpuamg", "VanekPC $theta 0.08 $vectLimit 60"], 1)):
                try:
                    line[key] = equal_size_inside_ug(0, 5, exe_name, ga, bsz)
                except Exception as e:
                    line[key] = {"error": str(e)[:200]}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=4, help="base cells per direction (N >= 2 with the defaults: SURVEY.md C3, 8x8x6)")
    ap.add_argument("--top", type=int, default=7, help="number of uniform refinements (levels - 1)")
    ap.add_argument("--weak", action="store_true", help="N >= 2: every GPU gets a box of cells^3 base cells instead of the fixed C3 problem")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-refine", type=int, default=6, help="refinements of the host-side reference run (6 -> 274 625 unknowns)")
    ap.add_argument("--cpu-replicas", type=int, default=0, help="concurrent single-threaded reference replicas (0 = one per host core, memory permitting)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline workload (no general-path / Q1 / elasticity / Galerkin / Krylov / weak lines)")
    ap.add_argument("--no-parity", action="store_true", help="N >= 2: skip the parity check against one GPU")
    ap.add_argument("--kind", default="p1", choices=["p1", "q1", "elasticity", "p1var"],
                    help="p1: BASELINE configs[1] (default); q1 / elasticity: Q1 cubes, scalar / 3x3 blocks (configs[3]; use --top 6); p1var: varying coefficient")
    ap.add_argument("--smoother", default="jac", choices=["jac", "gs", "sgs", "sor", "ilu"],
                    help="smoother class of the cycle: jac = BASELINE configs (default); gs / sgs / sor / ilu: Gauss-Seidel family and ILU (SURVEY.md 8f.2, one GPU)")
    ap.add_argument("--galerkin", action="store_true",
                    help="replace the coarse-level matrices by the Galerkin products P^T A P (uggpu_galerkin, SURVEY.md 8f.3) before the cycle "
                         "and report the time per level in config.galerkin_ms_top_down")
    ap.add_argument("--krylov", action="store_true", help="also time cg and bcgs around the cycle on the headline workload")
    ap.add_argument("--replicate-below", type=int, default=300000,
                    help="multi-GPU: levels with at most this many rows are held completely by every rank (coarse-level gather)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return our_arm(args)


if __name__ == "__main__":
    sys.exit(main())
