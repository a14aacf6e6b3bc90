"""Caller-supplied partition on real GPUs, run as  torchrun --nproc-per-node N tests/part_check.py  (N = 2 or 4).

A UG golden hierarchy (unstructured tetrahedra; adaptively refined tetrahedra) is partitioned by ug_b200/partition.py with the
reference's rules (RCB of the element centres, inheritance, lowest rank owns), every rank uploads its part through
uggpu_level_set_partition + the ordinary upload calls and all ranks solve with the V(2,2) cycle.  The gathered iterate and defect must
equal, BIT FOR BIT, what the unmodified reference left in its VECTORs after the same number of cycles (the dump's solve records), the
defect history to 1e-12.  Prints one line `PART-CHECK PASS ...` / `FAIL` per case (rank 0)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ug_b200 import capi, mgpu, partition  # noqa: E402
from ug_b200.hierarchy import Hierarchy  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
ARR = {2: (2, 1), 4: (2, 2), 8: (4, 2)}


def run(name, fused, repl, rank, world, local):
    hier = Hierarchy.from_ugh(os.path.join(GOLD, name + ".ugh"))
    d = hier.raw
    top, bs = hier.top, hier.bs
    dimx, dimy = ARR[world]
    owners = partition.vector_owners(hier, dimx, dimy)
    mine = partition.split(hier, owners, world, rank, repl)
    ctx = capi.Context(local)
    mgpu.init_comm(ctx, rank, world)
    ctx.upload_local_levels(mine, hier.fullrefinelevel, bs)
    A = ctx.handle("A")
    for l, L in enumerate(mine):
        for nm in ("x", "b", "c"):
            ctx.alloc(l, nm)
        ctx.put(l, "b", L.rhs)
    cycles = int(d["solve/cycles"][0])
    cfg = ctx.lmgc_cfg(nu1=int(hier.meta["nu1"]), nu2=int(hier.meta["nu2"]), gamma=int(hier.meta["gamma"]), baselevel=0, smooth_damp=hier.meta["damp"], fused=fused)
    ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
    res = capi.LResult()
    ctx.call("uggpu_ls_defect", 0, top, ctx.handle("x"), ctx.handle("b"), A)
    ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
    hist = np.zeros(cycles * bs)
    ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), cycles,
             capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), hist.ctypes.data_as(C.POINTER(C.c_double)))
    ok, worst = True, 0.0
    npart = sum(1 for L in mine if L.partitioned)
    exch = int(ctx.L.uggpu_comm_exchanges(ctx.h))
    for l in range(top + 1):
        L = mine[l]
        for nm, key in (("x", f"L{l}/solve/x_after_{cycles}"), ("b", f"L{l}/solve/b_after_{cycles}")):
            if key not in d:
                continue
            got = ctx.get(l, nm).reshape(-1, bs)
            want = d[key].reshape(-1, bs)[L.rows[:L.n]]
            ok = ok and np.array_equal(got, want)
    ctx.close()
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    href = d["solve/history"][:cycles * bs] if "solve/history" in d else None
    herr = float(np.max(np.abs(hist - href) / href)) if href is not None else -1.0
    good = bool(t.item() == 1.0) and (href is None or herr <= 1e-12)
    if rank == 0:
        print(f"PART-CHECK {'PASS' if good else 'FAIL'} " + json.dumps({"hierarchy": name, "ranks": world, "array": [dimx, dimy], "fused": fused, "levels": top + 1,
              "partitioned_levels": npart, "rows_rank0": [int(L.n) for L in mine], "ghosts_rank0": [int(L.n_ghost) for L in mine], "halo_exchanges": exch,
              "vectors_bitexact_vs_reference_dump": bool(t.item() == 1.0), "hist_relerr": herr, "cycles": cycles}), flush=True)
    return good


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def run_collect(name, repl, rank, world, local):
    """uggpu_l_vector_collect / uggpu_l_vector_consistent (l_vector_collect np/algebra/ugblas.cc:1035, l_vector_consistent :398): every rank
    holds ADDITIVE values -- its own part a_q[g] on every copy (owned or ghost) of vector g it holds; values are multiples of 2^-10, so the
    sums are exact in any order.  After collect the master copy must hold the sum over all ranks that hold a copy and the ghost rows 0;
    after consistent every copy must hold the sum."""
    hier = Hierarchy.from_ugh(os.path.join(GOLD, name + ".ugh"))
    top, bs = hier.top, hier.bs
    dimx, dimy = ARR[world]
    owners = partition.vector_owners(hier, dimx, dimy)
    parts = [partition.split(hier, owners, world, q, repl) for q in range(world)]
    mine = parts[rank]
    ctx = capi.Context(local)
    mgpu.init_comm(ctx, rank, world)
    ctx.upload_local_levels(mine, hier.fullrefinelevel, bs)
    ok, checked = True, 0
    for l, L in enumerate(mine):
        if not L.partitioned:
            continue
        ng = hier.levels[l].n
        part_of = lambda q: (np.round(np.random.default_rng(100 * l + q).standard_normal(ng * bs) * 1024) / 1024).reshape(ng, bs)
        want = np.zeros((ng, bs))
        for q in range(world):                       # the sum over the ranks that hold a copy of the vector
            held = np.zeros(ng, bool); held[parts[q][l].rows] = True
            want[held] += part_of(q)[held]
        for fn in ("uggpu_l_vector_collect", "uggpu_l_vector_consistent"):
            ctx.alloc(l, "v")
            t = torch.as_tensor(_DevArray(ctx.devptr(l, "v"), (L.n + L.n_ghost) * bs), device="cuda")
            t.copy_(torch.from_numpy(part_of(rank)[L.rows].reshape(-1)).cuda())
            torch.cuda.synchronize()
            ctx.call(fn, l, ctx.handle("v"))
            ctx.sync()
            got = t.cpu().numpy().reshape(-1, bs)
            ok = ok and np.array_equal(got[:L.n], want[L.rows[:L.n]])
            ghost_want = want[L.rows[L.n:]] if fn.endswith("consistent") else np.zeros((L.n_ghost, bs))
            ok = ok and np.array_equal(got[L.n:], ghost_want)
            checked += 1
    ctx.close()
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    good = bool(t.item() == 1.0) and checked > 0
    if rank == 0:
        print(f"PART-CHECK {'PASS' if good else 'FAIL'} " + json.dumps({"hierarchy": name, "ranks": world, "what": "l_vector_collect / l_vector_consistent on additive vectors",
              "levels_checked": checked // 2, "ghosts_rank0": [int(L.n_ghost) for L in mine]}), flush=True)
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, fused, repl in (("part_tet3d_r3", 1, 30), ("part_tet3d_r3", 0, 130), ("part_tet3d_adapt", 1, 30), ("part_tet3d_adapt", 0, 30)):
        ok = run(name, fused, repl, rank, world, local) and ok
    ok = run_collect("part_hex3d_bs3_r3", 30, rank, world, local) and ok
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
