"""Host-side logic of the multi-GPU path (ug_b200/csrc/part.h) on CPU: the node partition, the ghost numbering and the
halo message layout.  (1) exhaustive consistency of every rank pair for 2/4/8-rank arrays; (2) a real two-process
exchange over torch.distributed/gloo that follows the send/recv lists exactly as comm.cu does over NCCL."""
import ctypes as C
import itertools
import os
import socket

import numpy as np
import pytest

from ug_b200 import capi


def _lib():
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return capi.lib()


def describe(dim, cells, P, rank):
    L = _lib()
    out = (C.c_int32 * (16 + 16 * 26))()
    rc = L.uggpu_part_describe(dim, cells[0], cells[1], cells[2], P[0], P[1], P[2], rank, out, len(out))
    assert rc == 0, L.uggpu_last_error()
    o = np.array(out[:], dtype=np.int64)
    d = {"own": (o[0:3].copy(), o[3:6].copy()), "n_own": int(o[6]), "n_ghost": int(o[7]), "pitch": (int(o[9]), int(o[10])), "nb": []}
    for k in range(int(o[8])):
        q = o[16 + 16 * k: 32 + 16 * k]
        d["nb"].append({"rank": int(q[0]), "ns": int(q[1]), "nr": int(q[2]), "send": (q[3:6].copy(), q[6:9].copy()),
                        "recv": (q[9:12].copy(), q[12:15].copy())})
    return d


def box_nodes(box):
    lo, hi = box
    # lexicographic, x fastest (the order part.h enumerates a box in)
    return [(x, y, z) for z in range(lo[2], hi[2]) for y in range(lo[1], hi[1]) for x in range(lo[0], hi[0])]


def local_index(dim, cells, P, rank, x):
    return _lib().uggpu_part_local_index(dim, cells[0], cells[1], cells[2], P[0], P[1], P[2], rank, int(x[0]), int(x[1]), int(x[2]))


@pytest.mark.parametrize("dim,cells,P", [(3, (8, 4, 4), (2, 1, 1)), (3, (8, 8, 4), (2, 2, 1)), (3, (4, 4, 4), (2, 2, 2)),
                                         (2, (8, 8, 0), (2, 2, 1)), (3, (6, 6, 6), (3, 2, 1))])
def test_partition_consistency(dim, cells, P):
    nr = P[0] * P[1] * P[2]
    nn = [cells[d] + 1 if d < dim else 1 for d in range(3)]
    descs = [describe(dim, cells, P, r) for r in range(nr)]
    # owned boxes tile the grid
    owner = -np.ones(nn[::-1], dtype=int)
    for r, d in enumerate(descs):
        lo, hi = d["own"]
        assert np.all(owner[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] == -1)
        owner[lo[2]:hi[2], lo[1]:hi[1], lo[0]:hi[0]] = r
        # the owned box is numbered with odd pitches in all but the slowest direction (part.h: no power-of-two strides); the gaps are dummy rows
        ext = hi - lo
        want = [int(ext[k]) + (1 if (k < dim - 1 and ext[k] > 1 and ext[k] % 2 == 0) else 0) for k in range(2)]
        assert list(d["pitch"]) == want
        assert d["n_own"] == want[0] * want[1] * ext[2]
    assert np.all(owner >= 0)
    # a node on a cut plane belongs to the lower rank (priority.cc:200-222: master = lowest rank)
    if P[0] > 1:
        cut = cells[0] // P[0]
        assert np.all(owner[:, :, cut] < owner[:, :, cut + 1])
    for r, d in enumerate(descs):
        assert [nb["rank"] for nb in d["nb"]] == sorted(nb["rank"] for nb in d["nb"])
        for nb in d["nb"]:
            q = nb["rank"]
            back = [m for m in descs[q]["nb"] if m["rank"] == r]
            assert len(back) == 1
            # what r sends to q is, box for box and in the same order, what q receives from r
            assert all(np.array_equal(a, b) for a, b in zip(nb["send"], back[0]["recv"]))
            assert nb["ns"] == back[0]["nr"] == len(box_nodes(nb["send"]))
            for x in box_nodes(nb["recv"]):
                assert owner[x[2], x[1], x[0]] == q
        # local numbering: owned rows first (lexicographic in the owned box with its pitches), ghosts after, grouped by owner, gap-free
        seen = set()
        lo = d["own"][0]
        for x in box_nodes(d["own"]):
            i = (x[0] - lo[0]) + d["pitch"][0] * ((x[1] - lo[1]) + d["pitch"][1] * (x[2] - lo[2]))
            assert local_index(dim, cells, P, r, x) == i and i < d["n_own"]
            seen.add(i)
        off = d["n_own"]
        for nb in d["nb"]:
            for i, x in enumerate(box_nodes(nb["recv"])):
                assert local_index(dim, cells, P, r, x) == off + i
                seen.add(off + i)
            off += nb["nr"]
        assert off == d["n_own"] + d["n_ghost"] and len(seen) == len(box_nodes(d["own"])) + d["n_ghost"]
        # every stencil neighbour (Kuhn directions and their negatives) of an owned node is owned or a ghost
        dirs = [v for v in itertools.product((0, 1), repeat=3) if any(v) and (dim == 3 or v[2] == 0)]
        lo, hi = d["own"]
        for x in box_nodes(d["own"]):
            if not any(x[k] in (lo[k], hi[k] - 1) for k in range(dim)):
                continue
            for v in dirs:
                for s in (1, -1):
                    y = tuple(x[k] + s * v[k] for k in range(3))
                    if all(0 <= y[k] < nn[k] for k in range(3)):
                        assert local_index(dim, cells, P, r, y) >= 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, dim, cells, P, q):
    import torch
    import torch.distributed as dist
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        d = describe(dim, cells, P, rank)
        nn = [cells[k] + 1 if k < dim else 1 for k in range(3)]
        gid = lambda x: x[0] + nn[0] * (x[1] + nn[1] * x[2])
        v = torch.full((d["n_own"] + d["n_ghost"],), -1.0, dtype=torch.float64)
        own = box_nodes(d["own"])
        v[:d["n_own"]] = 0.0               # dummy rows of the padded numbering stay 0
        v[torch.tensor([local_index(dim, cells, P, rank, x) for x in own], dtype=torch.long)] = torch.tensor([float(gid(x)) for x in own], dtype=torch.float64)
        # pack (k_halo_pack), grouped send/recv into the ghost tail (comm.cu halo_exchange)
        reqs, off = [], d["n_own"]
        bufs = []
        for nb in d["nb"]:
            idx = torch.tensor([local_index(dim, cells, P, rank, x) for x in box_nodes(nb["send"])], dtype=torch.long)
            bufs.append(v[idx].contiguous())
            reqs.append(dist.isend(bufs[-1], nb["rank"]))
            reqs.append(dist.irecv(v[off:off + nb["nr"]], nb["rank"]))
            off += nb["nr"]
        for r in reqs:
            r.wait()
        # every ghost row now holds the value of its owner = its global id
        off = d["n_own"]
        ok = True
        for nb in d["nb"]:
            want = torch.tensor([float(gid(x)) for x in box_nodes(nb["recv"])], dtype=torch.float64)
            ok = ok and bool(torch.equal(v[off:off + nb["nr"]], want))
            off += nb["nr"]
        # the global sum of owner-masked local sums = the sequential sum (UG_GlobalSumNDOUBLE -> all_reduce)
        s = v[:d["n_own"]].sum().reshape(1)
        dist.all_reduce(s)
        n = nn[0] * nn[1] * nn[2]
        ok = ok and float(s) == n * (n - 1) / 2
        q.put((rank, ok))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.parametrize("dim,cells,P", [(3, (8, 4, 4), (2, 1, 1)), (2, (8, 8, 0), (1, 2, 1))])
def test_halo_exchange_two_ranks_gloo(dim, cells, P):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dim, cells, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
