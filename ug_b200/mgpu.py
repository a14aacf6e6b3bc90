"""Multi-GPU helpers shared by bench.py and tests/mgpu_check.py (one process per GPU under torchrun, NCCL).

`parity_check` is the driver-visible proof that a partitioned solve is the single-GPU solve: every rank builds its part
of a partitioned synthetic hierarchy and all ranks solve with the V(2,2) cycle; rank 0 repeats the solve unpartitioned
on its own GPU and compares.  Owner-computes rows with ghost COPIES evaluate every row with the same entries in the
same order as one GPU does, so the iterate x and the defect b must agree BIT FOR BIT; the defect history agrees to
1e-12 (the global sum is formed in another order).  The single-GPU path itself is pinned bit-exactly to the oracle
(tests/test_synth.py, tests/test_gpu_parity.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

ARRAYS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
KINDS = {"p1": capi.SYNTH_P1_SIMPLEX, "q1": capi.SYNTH_Q1_POISSON, "elasticity": capi.SYNTH_Q1_ELASTICITY, "p1var": capi.SYNTH_P1_VARCOEF}


def init_comm(ctx, rank, world):
    """ncclGetUniqueId on rank 0, broadcast through torch.distributed, uggpu_comm_init on every rank."""
    import torch
    import torch.distributed as dist
    idbuf = (C.c_char * 128)()
    if rank == 0:
        ctx.call_noctx("uggpu_comm_unique_id", idbuf)
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    ctx.call("uggpu_comm_init", world, rank, C.c_char_p(bytes(t.cpu().tolist())))


def _solve(ctx, top, cycles, fused):
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
    bs = ctx.level_bs(top)
    hist = np.zeros(cycles * bs)
    ctx.call("uggpu_ls_solve", C.byref(cfg), 0, top, ctx.handle("x"), ctx.handle("b"), A, ctx.handle("c"), cycles,
             capi._vs([1e-300]), capi._vs([1e-300]), C.byref(res), hist.ctypes.data_as(C.POINTER(C.c_double)))
    return np.array([res.first_defect[i] for i in range(bs)]), hist


def parity_check(rank, world, local, kind="p1", top=5, fused=1, base=2, cycles=6, replicate_below=5000, small_levels=False):
    """Returns a dict (the same on every rank) with ok, x_bitexact, b_bitexact, hist_relerr and what was run.
    small_levels: the two-kernel row forms (stx.cu, trc.cu: stencil / class rows + exception rows, the multi-GPU work in the exception
    kernels) also on levels below their size thresholds, so that a small test hierarchy takes the code paths of the bench."""
    import os
    import torch
    import torch.distributed as dist
    saved = {k: os.environ.get(k) for k in ("UGGPU_STX_MIN_ROWS", "UGGPU_TRC_MIN_ROWS")}
    if small_levels:
        os.environ["UGGPU_STX_MIN_ROWS"] = "2000"; os.environ["UGGPU_TRC_MIN_ROWS"] = "500"
    try:
        return _parity_check(rank, world, local, kind, top, fused, base, cycles, replicate_below, small_levels)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _parity_check(rank, world, local, kind, top, fused, base, cycles, replicate_below, small_levels):
    import os
    import torch
    import torch.distributed as dist
    P = ARRAYS[world]
    if os.environ.get("MGPU_ARRAY"):          # tests: another rank array for the same number of ranks, e.g. "1,2,1"
        P = tuple(int(v) for v in os.environ["MGPU_ARRAY"].split(","))
        assert P[0] * P[1] * P[2] == world
    cells = (base * P[0], base * P[1], base * P[2])
    ctx = capi.Context(local)
    init_comm(ctx, rank, world)
    ctx.call("uggpu_synth_hierarchy_part", KINDS[kind], cells[0], cells[1], cells[2], top, ctx.handle("A"),
             P[0], P[1], P[2], rank, C.c_int64(replicate_below))
    first, hist = _solve(ctx, top, cycles, fused)
    n, bs = ctx.level_n(top), ctx.level_bs(top)
    ids = np.zeros(n, np.int64)
    ctx.call("uggpu_synth_global_ids", top, ids.ctypes.data_as(C.c_void_p))
    x, b = ctx.get(top, "x"), ctx.get(top, "b")
    nparts = sum(int(ctx.L.uggpu_level_is_partitioned(ctx.h, l)) for l in range(top + 1))
    exch = int(ctx.L.uggpu_comm_exchanges(ctx.h))
    transport = ctx.halo_transport()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n], dtype=torch.int64, device="cuda"))
    nmax = int(max(s.item() for s in sizes))

    def gather(a, dtype, width):
        t = torch.zeros(nmax * width, dtype=dtype, device="cuda")
        t[:a.size] = torch.from_numpy(a).cuda()
        out = [torch.zeros(nmax * width, dtype=dtype, device="cuda") for _ in range(world)]
        dist.all_gather(out, t)
        return [o[:int(s.item()) * width].cpu().numpy() for o, s in zip(out, sizes)]

    ids_all, x_all, b_all = gather(ids, torch.int64, 1), gather(x, torch.float64, bs), gather(b, torch.float64, bs)
    ctx.close()
    out = torch.zeros(6, dtype=torch.float64, device="cuda")
    ng = (cells[0] * 2 ** top + 1) * (cells[1] * 2 ** top + 1) * (cells[2] * 2 ** top + 1)
    if rank == 0:
        xg, bg, cnt = np.zeros((ng, bs)), np.zeros((ng, bs)), np.zeros(ng, int)
        dummy_ok = True
        for i, xx, bb in zip(ids_all, x_all, b_all):
            real = i >= 0                  # ids -1: dummy rows of the padded local numbering (part.h): must stay 0
            xx, bb = xx.reshape(-1, bs), bb.reshape(-1, bs)
            dummy_ok = dummy_ok and not np.any(xx[~real]) and not np.any(bb[~real])
            xg[i[real]] = xx[real]; bg[i[real]] = bb[real]; np.add.at(cnt, i[real], 1)
        one = capi.Context(local)
        one.call("uggpu_synth_hierarchy", KINDS[kind], cells[0], cells[1], cells[2], top, one.handle("A"))
        first1, hist1 = _solve(one, top, cycles, fused)
        x1, b1 = one.get(top, "x").reshape(-1, bs), one.get(top, "b").reshape(-1, bs)
        one.close()
        xe, be = bool(np.array_equal(xg, x1)), bool(np.array_equal(bg, b1))
        herr = float(max(np.max(np.abs(hist - hist1) / hist1), np.max(np.abs(first - first1) / first1)))
        conv = bool(hist[-1] < 0.05 * hist[bs - 1])
        ok = bool(np.all(cnt == 1)) and dummy_ok and xe and be and herr <= 1e-12 and conv
        out = torch.tensor([1.0 if ok else 0.0, 1.0 if xe else 0.0, 1.0 if be else 0.0, herr, hist[bs - 1], hist[-1]], dtype=torch.float64, device="cuda")
    dist.broadcast(out, 0)
    o = out.cpu().tolist()
    return {"ok": o[0] == 1.0, "x_bitexact": o[1] == 1.0, "b_bitexact": o[2] == 1.0, "hist_relerr": o[3], "defect": [o[4], o[5]], "kind": kind,
            "ranks": world, "array": list(P), "fused": fused, "unknowns": ng * bs, "partitioned_levels": nparts, "levels": top + 1,
            "halo_exchanges": exch, "transport": transport, "cycles": cycles, "row_forms_on_small_levels": bool(small_levels),
            "against": "the same solve on ONE GPU (rank 0), itself pinned bit-exactly to the oracle by tests/test_synth.py"}
