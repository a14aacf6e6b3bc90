"""ug_b200/partition.py on CPU: the element partition of a UG golden hierarchy by the reference's rules (RCB of the level-0 element
centres parallel/dddif/lbrcb.cc:250-330, inheritance :376, lowest rank owns priority.cc:200-222) and the owner-computes split the
C-ABI's uggpu_level_set_partition takes.  Checks without a GPU: every vector has one owner; both sides of every interface list the
same vectors in the same order; with the ghost rows filled by the halo copy, every rank's rows of A y, P c and (summed over the ranks
across the gather level) R d are BIT-IDENTICAL to the unpartitioned products -- same entries, same order."""
import os

import numpy as np
import pytest

from ug_b200.hierarchy import Hierarchy
from ug_b200 import partition

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def seq_rows(rowptr, col, val, y, bs=1):
    """Row sums in stored order with separate multiply and add (the reference's arithmetic)."""
    n = rowptr.size - 1
    out = np.zeros(n * bs)
    v = val.reshape(-1, bs, bs) if bs > 1 else None
    for r in range(n):
        if bs == 1:
            s = 0.0
            for e in range(rowptr[r], rowptr[r + 1]):
                s += val[e] * y[col[e]]
            out[r] = s
        else:
            s = np.zeros(bs)
            for e in range(rowptr[r], rowptr[r + 1]):
                for i in range(bs):
                    acc = v[e, i, 0] * y[col[e] * bs]
                    for q in range(1, bs):
                        acc = acc + v[e, i, q] * y[col[e] * bs + q]
                    s[i] += acc
            out[r * bs:(r + 1) * bs] = s
    return out


@pytest.mark.parametrize("name,dimx,dimy,repl", [("part_tet3d_r3", 2, 1, 30), ("part_tet3d_r3", 2, 2, 130), ("part_tet3d_adapt", 2, 2, 30), ("part_tet3d_adapt", 4, 2, 30)])
def test_partition_of_a_ug_hierarchy(name, dimx, dimy, repl):
    hier = Hierarchy.from_ugh(os.path.join(GOLD, name + ".ugh"))
    nr = dimx * dimy
    owners = partition.vector_owners(hier, dimx, dimy)
    parts = [partition.split(hier, owners, nr, q, repl) for q in range(nr)]
    rng = np.random.default_rng(7)
    bs = hier.bs
    for l, lv in enumerate(hier.levels):
        locs = [p[l] for p in parts]
        y = np.round(rng.standard_normal(lv.n * bs) * 1024) / 1024
        ref = seq_rows(lv.rowptr, lv.col, lv.val, y, bs)
        if not locs[0].partitioned:
            assert all(not L.partitioned and L.n == lv.n for L in locs)
        else:
            # every vector has exactly one owner, ranks are balanced the way RCB balances elements (no empty rank on these grids)
            cnt = np.zeros(lv.n, int)
            for L in locs:
                cnt[L.rows[:L.n]] += 1
            assert np.all(cnt == 1) and ("adapt" in name or all(L.n > 0 for L in locs))      # locally refined levels may leave ranks empty
            for q, L in enumerate(locs):
                assert np.all(owners[l][L.rows[:L.n]] == q) and np.all(owners[l][L.rows[L.n:]] != q)
                # interfaces: what q sends to k is, vector for vector, what k receives from q
                for i, k in enumerate(L.nb_rank):
                    K = locs[k]
                    j = list(K.nb_rank).index(q)
                    sent = L.rows[L.send_idx[L.send_off[i]:L.send_off[i + 1]]]
                    got = K.rows[K.n + K.recv_off[j]:K.n + K.recv_off[j + 1]]
                    assert np.array_equal(sent, got)
                # A y on the rank's rows with ghost rows filled by the copy: bit-identical
                yl = y.reshape(-1, bs)[L.rows].reshape(-1)
                assert np.array_equal(seq_rows(L.rowptr, L.col, L.val, yl, bs), ref.reshape(-1, bs)[L.rows[:L.n]].reshape(-1))
        if l > 0 and bs == 1:
            cl = hier.levels[l - 1]
            c = np.round(rng.standard_normal(cl.n) * 1024) / 1024
            d = np.round(rng.standard_normal(lv.n) * 1024) / 1024
            pref = seq_rows(lv.p_rowptr, lv.p_col, lv.p_w, c)
            rref = seq_rows(lv.r_rowptr, lv.r_col, lv.r_w, d)
            racc = np.zeros(cl.n)
            for q, L in enumerate(locs):
                C = parts[q][l - 1]
                got = seq_rows(L.p_rowptr, L.p_col, L.p_w, c[C.rows])
                assert np.array_equal(got, pref[L.rows[:L.n]])
                rl = seq_rows(L.r_rowptr, L.r_col, L.r_w, d[L.rows])
                if C.partitioned:
                    assert np.array_equal(rl, rref[C.rows[:C.n]])
                elif L.partitioned:
                    racc += rl                   # the gather level: disjoint parts, added by the all-reduce
                else:
                    assert np.array_equal(rl, rref)
            if locs[0].partitioned and not parts[0][l - 1].partitioned:
                assert np.array_equal(racc, rref)


def test_rcb_follows_the_reference_on_a_regular_grid():
    """theRCB on a 4 x 4 (x 4) array of unit cells: halves in x, then y, ... with destination py * DimX + px (lbrcb.cc:283-319)."""
    xs = np.arange(4) + 0.5
    c2 = np.array([(x, y) for y in xs for x in xs])
    d = partition.rcb_elements(c2, 2, 2)
    want = np.array([(0 if x < 2 else 1) + 2 * (0 if y < 2 else 1) for y in xs for x in xs])
    assert np.array_equal(d, want)
    c3 = np.array([(x, y, z) for z in xs for y in xs for x in xs])
    d = partition.rcb_elements(c3, 4, 2)          # 8 ranks: x halved (sorted by x), x halved again (sorted by y!), then y (sorted by z)
    assert sorted(np.bincount(d).tolist()) == [8] * 8
    px, py = d % 4, d // 4
    assert np.all((px >= 2) == (c3[:, 0] > 2)) and np.all((px % 2 == 1) == (c3[:, 1] > 2)) and np.all((py == 1) == (c3[:, 2] > 2))
