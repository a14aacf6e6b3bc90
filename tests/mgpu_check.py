"""Multi-GPU parity check, run as  torchrun --nproc-per-node N tests/mgpu_check.py  (N = 2, 4 or 8 GPUs of one box).

Every rank builds its part of a partitioned synthetic hierarchy, all ranks solve with the fused V(2,2) cycle over NCCL;
rank 0 repeats the solve unpartitioned on its own GPU and compares: the iterate x and the defect b must agree BIT FOR
BIT (owner-computes with ghost copies evaluates every row exactly as one GPU does), the defect history to 1e-12
(the global sum is formed in a different order).  Prints one line `MGPU-CHECK PASS ...` / `FAIL`."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ug_b200 import capi  # noqa: E402

ARRAYS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def init_comm(ctx, rank, world):
    idbuf = (C.c_char * 128)()
    if rank == 0:
        rc = ctx.L.uggpu_comm_unique_id(idbuf)
        assert rc == 0, ctx.L.uggpu_last_error()
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    ctx.call("uggpu_comm_init", world, rank, C.c_char_p(raw))


def solve(ctx, top, cycles, fused):
    A = ctx.handle("A")
    for name in ("x", "b", "c"):
        for l in range(top + 1):
            ctx.alloc(l, name)
    ctx.call("uggpu_synth_rhs", top, ctx.handle("b"))
    cfg = ctx.lmgc_cfg(nu1=2, nu2=2, gamma=1, baselevel=0, smooth_damp=0.6, fused=fused)
    ctx.call("uggpu_lmgc_preprocess", C.byref(cfg), top, A)
    res = capi.LResult()
    ctx.call("uggpu_ls_defect", 0, top, ctx.handle("x"), ctx.handle("b"), A)
    ctx.call("uggpu_ls_residuum", 0, top, ctx.handle("b"), C.byref(res))
    hist = np.zeros(cycles)
    ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), cycles,
             capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), hist.ctypes.data_as(C.POINTER(C.c_double)))
    return res.first_defect[0], hist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = ARRAYS[world]
    base, top, cycles = 2, int(os.environ.get("MGPU_TOP", "5")), 6
    cells = (base * P[0], base * P[1], base * P[2])
    ok_all = True
    for fused in (1, 0):
        ctx = capi.Context(local)
        init_comm(ctx, rank, world)
        ctx.call("uggpu_synth_hierarchy_part", capi.SYNTH_P1_SIMPLEX, cells[0], cells[1], cells[2], top, ctx.handle("A"),
                 P[0], P[1], P[2], rank, C.c_int64(int(os.environ.get("MGPU_REPL", "5000"))))
        first, hist = solve(ctx, top, cycles, fused)
        n = ctx.level_n(top)
        ids = np.zeros(n, np.int64)
        ctx.call("uggpu_synth_global_ids", top, ids.ctypes.data_as(C.c_void_p))
        x, b = ctx.get(top, "x"), ctx.get(top, "b")
        nparts = sum(int(ctx.L.uggpu_level_is_partitioned(ctx.h, l)) for l in range(top + 1))
        exch = int(ctx.L.uggpu_comm_exchanges(ctx.h))
        # gather the parts on rank 0
        sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
        nmax = int(max(s.item() for s in sizes))
        def gather(a, dtype):
            t = torch.zeros(nmax, dtype=dtype, device="cuda")
            t[:n] = torch.from_numpy(a).cuda()
            out = [torch.zeros(nmax, dtype=dtype, device="cuda") for _ in range(world)]
            dist.all_gather(out, t)
            return [o[:int(s.item())].cpu().numpy() for o, s in zip(out, sizes)]
        ids_all, x_all, b_all = gather(ids, torch.int64), gather(x, torch.float64), gather(b, torch.float64)
        ctx.close()
        if rank == 0:
            ng = (cells[0] * 2 ** top + 1) * (cells[1] * 2 ** top + 1) * (cells[2] * 2 ** top + 1)
            xg, bg, cnt = np.zeros(ng), np.zeros(ng), np.zeros(ng, int)
            for i, xx, bb in zip(ids_all, x_all, b_all):
                xg[i] = xx; bg[i] = bb; cnt[i] += 1
            one = capi.Context(local)
            one.call("uggpu_synth_hierarchy", capi.SYNTH_P1_SIMPLEX, cells[0], cells[1], cells[2], top, one.handle("A"))
            first1, hist1 = solve(one, top, cycles, fused)
            x1, b1 = one.get(top, "x"), one.get(top, "b")
            one.close()
            ok = (np.all(cnt == 1) and np.array_equal(xg, x1) and np.array_equal(bg, b1)
                  and abs(first - first1) <= 1e-12 * first1 and np.max(np.abs(hist - hist1) / hist1) <= 1e-12 and hist[-1] < 0.05 * hist[0])
            print(f"MGPU-CHECK {'PASS' if ok else 'FAIL'} ranks={world} array={P} fused={fused} unknowns={ng} partitioned_levels={nparts}/{top + 1} "
                  f"halo_exchanges={exch} x_bitexact={np.array_equal(xg, x1)} b_bitexact={np.array_equal(bg, b1)} "
                  f"hist_relerr={np.max(np.abs(hist - hist1) / hist1):.2e} defect {hist[0]:.3e}->{hist[-1]:.3e}", flush=True)
            ok_all = ok_all and ok
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
