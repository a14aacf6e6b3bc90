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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name, fused, repl in (("part_tet3d_r3", 1, 30), ("part_tet3d_r3", 0, 130), ("part_tet3d_adapt", 1, 30), ("part_tet3d_adapt", 0, 30)):
        ok = run(name, fused, repl, rank, world, local) and ok
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
