"""Element partition of a UG hierarchy over several GPUs, the way UG's own parallel data model does it (host side, numpy).

This is what a ModelP caller of the `gpuls` numprocs hands to the device layer (C-ABI `uggpu_level_set_partition`); it restates, for a
hierarchy held on one host, the three rules of the reference that decide which rank computes what:

  * `rcb_elements`  -- the load balancer of parallel/dddif/lbrcb.cc:250-330 (`theRCB`): recursive coordinate bisection of the centres of
    mass (`CenterOfMass` :346) of the LEVEL-0 elements over a DimX x DimY processor array, sort direction cycling x, y, z, the longer
    side of the processor array halved, `ni0 = (int)(part0 / d * nItems)`; ties broken by the other coordinates like `sort_rcb_x/y/z`
    (:100-222); destination `py * DimX + px`;
  * `InheritPartition` (:376): sons take their father's partition, so the element partitions of all levels are nested;
  * `ComputeVectorBorderPrios` (parallel/dddif/priority.cc:200-222): a vector shared by several partitions is master on the LOWEST rank.

On top of that the device layer's storage model: owner computes -- a rank holds the FULL rows of the vectors it owns and
ghost COLUMNS for everything those rows (and its rows of P and R) reference, so that every row is evaluated with the same entries in the
same order as on one GPU and a halo COPY of the operand replaces UG's halo SUM of the result (l_vector_consistent,
np/algebra/ugblas.cc:398).  Levels with at most `replicate_below` vectors are held completely by every rank (coarse-level
agglomeration, np/procs/amgtransfer.cc:246 `$aggLimit`); the restriction into the first such level fills, on every rank, only the rows
of the coarse vectors that rank owns, and the all-reduce that follows adds the disjoint parts.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field
from typing import List

import numpy as np

SMALL_DOUBLE = 1e-20       # low/misc.h SMALL_D is DBL_EPSILON*10; lbrcb.cc compares with SMALL_DOUBLE; any tiny tolerance orders the same centres


def _cmp_factory(order):
    def cmp(a, b):
        for d in order:
            if a[1][d] < b[1][d] - SMALL_DOUBLE:
                return -1
            if a[1][d] > b[1][d] + SMALL_DOUBLE:
                return 1
        return 0
    return cmp


def rcb_elements(centers: np.ndarray, dimx: int, dimy: int) -> np.ndarray:
    """parallel/dddif/lbrcb.cc theRCB: destination rank of every element (centers: [ne, dim])."""
    ne, dim = centers.shape
    dest = np.zeros(ne, np.int32)
    # comparison orders of sort_rcb_x / _y / _z (lbrcb.cc:100-222): primary coordinate, then the others as the reference lists them
    orders = {0: (0, 1, 2)[:dim], 1: (1, 0, 2)[:dim], 2: (2, 1, 0)[:dim]}

    def rec(items, px, py, dx, dy, d):
        if not items:
            return
        if dx <= 1 and dy <= 1:
            for i, _ in items:
                dest[i] = py * dimx + px
            return
        if len(items) > 1:
            items = sorted(items, key=functools.cmp_to_key(_cmp_factory(orders[d])))
        if dx >= dy:
            p0 = dx // 2
            n0 = int(float(p0) / float(dx) * float(len(items)))
            rec(items[:n0], px, py, p0, dy, (d + 1) % dim)
            rec(items[n0:], px + p0, py, dx - p0, dy, (d + 1) % dim)
        else:
            p0 = dy // 2
            n0 = int(float(p0) / float(dy) * float(len(items)))
            rec(items[:n0], px, py, dx, p0, (d + 1) % dim)
            rec(items[n0:], px, py + p0, dx, dy - p0, (d + 1) % dim)

    rec([(i, centers[i]) for i in range(ne)], 0, 0, dimx, dimy, 0)
    return dest


def vector_owners(hier, dimx: int, dimy: int) -> List[np.ndarray]:
    """Owner rank of every vector of every level: RCB of the level-0 elements, inheritance, lowest rank among the elements at a vector."""
    d = hier.raw
    dim = hier.dim
    owners, elem_rank_below = [], None
    for l, lv in enumerate(hier.levels):
        ep, en, ef = d[f"L{l}/elem_ptr"], d[f"L{l}/elem_nodes"], d[f"L{l}/elem_father"]
        xyz = lv.xyz.reshape(-1, dim)
        ne = ep.size - 1
        if l == 0:
            centers = np.stack([xyz[en[ep[e]:ep[e + 1]]].sum(0) * (1.0 / float(ep[e + 1] - ep[e])) for e in range(ne)])
            er = rcb_elements(centers, dimx, dimy)
        else:
            er = elem_rank_below[ef]                                  # InheritPartition
        own = np.full(lv.n, np.iinfo(np.int32).max, np.int32)
        for e in range(ne):
            nodes = en[ep[e]:ep[e + 1]]
            own[nodes] = np.minimum(own[nodes], er[e])                # priority.cc:200: master = lowest rank
        assert own.max() < dimx * dimy, "a vector without an element on its level"
        owners.append(own)
        elem_rank_below = er
    return owners


@dataclass
class LocalLevel:
    """What one rank holds of one level (the arguments of uggpu_level_create / _set_partition / _set_flags / uggpu_mat_set / uggpu_transfer_set)."""
    partitioned: bool
    n: int                                  # rows this rank holds (owned rows; all rows on a level held completely)
    n_ghost: int
    n_global: int
    rows: np.ndarray                        # global row of every local row (owned, then ghosts)
    nb_rank: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    send_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    send_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_off: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    rowptr: np.ndarray = None
    col: np.ndarray = None
    val: np.ndarray = None
    vclass: np.ndarray = None
    vnclass: np.ndarray = None
    ctl: np.ndarray = None
    skip: np.ndarray = None
    p_rowptr: np.ndarray = None
    p_col: np.ndarray = None
    p_w: np.ndarray = None
    r_rowptr: np.ndarray = None
    r_col: np.ndarray = None
    r_w: np.ndarray = None
    rhs: np.ndarray = None


def _sub_csr(rowptr, col, w, rows, colmap, bb=1):
    """Rows `rows` of a CSR matrix, entries in their stored order, columns through colmap (global -> local; must be >= 0)."""
    cnt = rowptr[rows + 1] - rowptr[rows]
    rp = np.zeros(rows.size + 1, np.int32)
    np.cumsum(cnt, out=rp[1:])
    idx = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in rows]) if rows.size else np.zeros(0, np.int64)
    c = colmap[col[idx]]
    assert c.size == 0 or c.min() >= 0, "a referenced column is neither owned nor a ghost"
    ww = w.reshape(-1, bb)[idx].reshape(-1) if w is not None else None
    return rp, c.astype(np.int32), ww


def split(hier, owners: List[np.ndarray], nranks: int, rank: int, replicate_below: int) -> List[LocalLevel]:
    """The part of the hierarchy rank `rank` holds.  Ghost sets of a level: the columns of the rank's rows of A, of its rows of P on the
    level above, and of its rows of R on the level below, that other ranks own; grouped by owner, ascending global row inside a group --
    the same enumeration on both sides of an interface (the role of the sorted interface lists of parallel/ddd/if/ifcreate.cc:155-203)."""
    top = len(hier.levels) - 1
    repl = [lv.n <= replicate_below or l == 0 for l, lv in enumerate(hier.levels)]
    bs = hier.bs
    bb = bs * bs

    def own_rows(l, q):
        return np.nonzero(owners[l] == q)[0]

    def needed(l, q):
        """Global rows of level l that rank q reads without owning them."""
        lv = hier.levels[l]
        mine = own_rows(l, q)
        need = [lv.col[np.concatenate([np.arange(lv.rowptr[r], lv.rowptr[r + 1]) for r in mine])]] if mine.size else []
        if l < top and not repl[l + 1]:      # P rows of my fine vectors read coarse values of this level
            f = hier.levels[l + 1]
            fm = own_rows(l + 1, q)
            if fm.size:
                need.append(f.p_col[np.concatenate([np.arange(f.p_rowptr[r], f.p_rowptr[r + 1]) for r in fm])])
        if l > 0:                            # R rows of my coarse vectors read fine values of this level
            cm = own_rows(l - 1, q)
            if cm.size:
                need.append(lv.r_col[np.concatenate([np.arange(lv.r_rowptr[r], lv.r_rowptr[r + 1]) for r in cm])])
        if not need:
            return np.zeros(0, np.int64)
        allc = np.unique(np.concatenate(need))
        return allc[owners[l][allc] != q]

    out = []
    colmaps = []
    for l, lv in enumerate(hier.levels):
        if repl[l]:
            rows = np.arange(lv.n)
            L = LocalLevel(partitioned=False, n=lv.n, n_ghost=0, n_global=lv.n, rows=rows)
            cmap = np.arange(lv.n)
        else:
            mine = own_rows(l, rank)
            gh = needed(l, rank)
            gh = gh[np.lexsort((gh, owners[l][gh]))]                 # by owner, then ascending global row
            rows = np.concatenate([mine, gh])
            cmap = np.full(lv.n, -1, np.int64)
            cmap[rows] = np.arange(rows.size)
            nb, soff, sidx, roff = [], [0], [], [0]
            for q in range(nranks):
                if q == rank:
                    continue
                recv = gh[owners[l][gh] == q]
                theirs = needed(l, q)
                send = theirs[owners[l][theirs] == rank]              # ascending global row: q's ghost order for my rows
                if recv.size == 0 and send.size == 0:
                    continue
                nb.append(q)
                sidx.append(cmap[send])
                soff.append(soff[-1] + send.size)
                roff.append(roff[-1] + recv.size)
            L = LocalLevel(partitioned=True, n=mine.size, n_ghost=gh.size, n_global=lv.n, rows=rows, nb_rank=np.array(nb, np.int32),
                           send_off=np.array(soff, np.int32), send_idx=(np.concatenate(sidx) if sidx else np.zeros(0)).astype(np.int32),
                           recv_off=np.array(roff, np.int32))
        own = rows[:L.n]
        L.rowptr, L.col, L.val = _sub_csr(lv.rowptr, lv.col, lv.val, own, cmap, bb)
        L.vclass, L.vnclass, L.ctl, L.skip = lv.vclass[own].copy(), lv.vnclass[own].copy(), lv.ctl[own].copy(), lv.skip[own].copy()
        L.rhs = lv.rhs.reshape(-1, bs)[own].reshape(-1).copy() if lv.rhs is not None else None
        if l > 0:
            cl, cmap_c = out[l - 1], colmaps[l - 1]
            # P: my fine rows -> coarse columns (owned, ghost, or any row of a completely held coarse level)
            L.p_rowptr, L.p_col, L.p_w = _sub_csr(lv.p_rowptr, lv.p_col, lv.p_w, own, cmap_c)
            # R: rows = the coarse rows I hold; on a completely held coarse level below a partitioned fine one only the coarse vectors I own are
            # filled (the all-reduce adds the ranks' disjoint parts), the other rows are empty
            crow = cl.rows[:cl.n]
            if repl[l - 1] and not repl[l]:
                assert np.all(hier.levels[l - 1].vnclass >= 2), "the gather level must have VNCLASS >= NEWDEF_CLASS everywhere (rows start at 0 on every rank)"
                keep = owners[l - 1][crow] == rank
                rp = np.zeros(crow.size + 1, np.int32)
                cnt = (lv.r_rowptr[crow + 1] - lv.r_rowptr[crow]) * keep
                np.cumsum(cnt, out=rp[1:])
                sel = crow[keep]
                _, cc, ww = _sub_csr(lv.r_rowptr, lv.r_col, lv.r_w, sel, cmap)
                L.r_rowptr, L.r_col, L.r_w = rp, cc, ww
            else:
                L.r_rowptr, L.r_col, L.r_w = _sub_csr(lv.r_rowptr, lv.r_col, lv.r_w, crow, cmap)
        out.append(L)
        colmaps.append(cmap)
    return out
