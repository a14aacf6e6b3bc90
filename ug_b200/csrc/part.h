// part.h -- node partition of the structured synthetic grids over a Px x Py x Pz array of GPUs (host + device).
//
// Follows the reference's parallel data model where it matters for results:
//  * the ELEMENT partition is a recursive coordinate bisection into equal boxes of base-level cells
//    (parallel/dddif/lbrcb.cc:250-330 on a structured grid yields exactly such boxes); sons inherit the father's
//    partition (lbrcb.cc:376), so the partitions of all levels are nested;
//  * a vector shared by several partitions is master on the LOWEST rank (parallel/dddif/priority.cc:200-222): with
//    ranks numbered x-fastest the owner of a node on a cut plane is the lower box;
//  * instead of UG's additive storage + sum exchange (l_vector_consistent, np/algebra/ugblas.cc:398) every rank holds
//    the FULL rows of the vectors it owns plus one layer of ghost columns (owner computes).  A halo COPY of the
//    operand replaces the halo SUM of the result, and every row is evaluated with the same entries in the same order
//    as on one GPU: partitioned results are bit-identical to unpartitioned ones.
//
// Local numbering on a partitioned level: owned nodes first (lexicographic inside the owned box), then the ghost
// nodes grouped by owner rank (ascending), lexicographic inside (owned box of the owner) n (my box grown by one).
// The owned box is numbered with ODD pitches in x and y: an owned box of 2^k cells has 2^k nodes per direction on every rank but
// the first (the node on the cut plane belongs to the lower rank), and with a power-of-two pitch the 15 / 27 gathers of a stencil
// row land in the same few L1 sets -- measured on B200: the stencil kernel of the upper half of an x-split (512-wide box) ran 10 %
// slower than the lower half's (513-wide).  An even extent e is therefore numbered with pitch e + 1; the extra indices are DUMMY
// rows: empty matrix rows, VCLASS 0, no flags, all skip bits, never a column -- inert in every kernel, 0.2-0.4 % of the rows.
// Both sides of an interface enumerate the same box in the same order, which fixes the message layout (the role of
// the sorted interface lists of parallel/ddd/if/ifcreate.cc:155-203).
#ifndef UGGPU_PART_H
#define UGGPU_PART_H

#include <stdint.h>

#ifdef __CUDACC__
#define PART_HD __host__ __device__ __forceinline__
#else
#define PART_HD inline
#endif

#define PART_MAX_NB 26

struct PartBox { int lo[3], hi[3]; };   // half-open node index ranges

struct PartGrid {
  int dim;
  int P[3];            // GPU array
  int coord[3];        // my position
  int rank, nranks;
  int replicated;      // 1: this level is held completely (global lexicographic numbering) by every rank
  int nn[3];           // global nodes per direction on this level
  int cpr[3];          // cells per rank and direction on this level (level cells / P)
  int pitch[2];        // row-numbering pitches of the owned box in x and y (>= its extents; see above)
  PartBox own;         // nodes I own (replicated level: the nodes I would own, used for the restriction rows)
  PartBox ext;         // own grown by one layer, clipped to the domain
  int n_own, n_ghost;
  int nnb;
  int nb_rank[PART_MAX_NB];
  int nb_slot[27];     // (dx+1) + 3(dy+1) + 9(dz+1) -> neighbour index or -1
  PartBox nb_recv[PART_MAX_NB];   // owned(nb) n ext(me): my ghosts owned by nb
  int nb_recv_off[PART_MAX_NB + 1];
  PartBox nb_send[PART_MAX_NB];   // owned(me) n ext(nb): what nb needs from me
  int nb_send_off[PART_MAX_NB + 1];
};

PART_HD int box_count(const PartBox &b)
{
  int c = 1;
  for (int d = 0; d < 3; d++) { int e = b.hi[d] - b.lo[d]; if (e <= 0) return 0; c *= e; }
  return c;
}
PART_HD bool box_has(const PartBox &b, const int x[3])
{
  return x[0] >= b.lo[0] && x[0] < b.hi[0] && x[1] >= b.lo[1] && x[1] < b.hi[1] && x[2] >= b.lo[2] && x[2] < b.hi[2];
}
PART_HD int box_lex(const PartBox &b, const int x[3])
{
  return (x[0] - b.lo[0]) + (b.hi[0] - b.lo[0]) * ((x[1] - b.lo[1]) + (b.hi[1] - b.lo[1]) * (x[2] - b.lo[2]));
}
PART_HD void box_unlex(const PartBox &b, int i, int x[3])
{
  int e0 = b.hi[0] - b.lo[0], e1 = b.hi[1] - b.lo[1];
  x[0] = b.lo[0] + i % e0;
  int q = i / e0;
  x[1] = b.lo[1] + q % e1;
  x[2] = b.lo[2] + q / e1;
}
PART_HD PartBox box_and(const PartBox &a, const PartBox &b)
{
  PartBox r;
  for (int d = 0; d < 3; d++) { r.lo[d] = a.lo[d] > b.lo[d] ? a.lo[d] : b.lo[d]; r.hi[d] = a.hi[d] < b.hi[d] ? a.hi[d] : b.hi[d]; }
  return r;
}

// owned node range of position a along a direction with `cpr` cells per rank: the node on the lower cut plane
// belongs to the lower neighbour
PART_HD void part_own_range(int a, int cpr, int *lo, int *hi) { *lo = a * cpr + (a > 0 ? 1 : 0); *hi = (a + 1) * cpr + 1; }

// position of the owner of node coordinate x
PART_HD int part_owner_coord(int x, int cpr, int P) { int a = x == 0 ? 0 : (x - 1) / cpr; return a < P ? a : P - 1; }

// index of owned node x in the padded numbering of the owned box
PART_HD int part_own_lex(const PartGrid &g, const int x[3])
{
  return (x[0] - g.own.lo[0]) + g.pitch[0] * ((x[1] - g.own.lo[1]) + g.pitch[1] * (x[2] - g.own.lo[2]));
}

// local index of global node x on this level (x must be owned or a ghost of this rank); -1 otherwise
PART_HD int part_local_index(const PartGrid &g, const int x[3])
{
  if (g.replicated) return x[0] + g.nn[0] * (x[1] + g.nn[1] * x[2]);
  if (box_has(g.own, x)) return part_own_lex(g, x);
  int slot = 0, mul = 1;
  for (int d = 0; d < 3; d++) {
    int da = (d < g.dim ? part_owner_coord(x[d], g.cpr[d], g.P[d]) : 0) - g.coord[d];
    if (da < -1 || da > 1) return -1;
    slot += (da + 1) * mul;
    mul *= 3;
  }
  int k = g.nb_slot[slot];
  if (k < 0 || !box_has(g.nb_recv[k], x)) return -1;
  return g.n_own + g.nb_recv_off[k] + box_lex(g.nb_recv[k], x);
}

// global coordinates of local row r (owned rows only on partitioned levels); false: r is a dummy row of the padded numbering
// (x then holds the coordinates of a real node of the box, so that callers which ignore the result stay in range)
PART_HD bool part_row_coords(const PartGrid &g, int r, int x[3])
{
  if (g.replicated) { x[0] = r % g.nn[0]; int q = r / g.nn[0]; x[1] = q % g.nn[1]; x[2] = q / g.nn[1]; return true; }
  int a = r % g.pitch[0], q = r / g.pitch[0];
  int b = q % g.pitch[1], c = q / g.pitch[1];
  const int e0 = g.own.hi[0] - g.own.lo[0], e1 = g.own.hi[1] - g.own.lo[1];
  const bool real = a < e0 && b < e1;
  if (a >= e0) a = e0 - 1;
  if (b >= e1) b = e1 - 1;
  x[0] = g.own.lo[0] + a; x[1] = g.own.lo[1] + b; x[2] = g.own.lo[2] + c;
  return real;
}

// Fills g for `rank` of a P[0] x P[1] x P[2] array; cells[d] = cells of this level in direction d (0 beyond dim).
// Returns non-zero if the cells do not divide evenly.
inline int part_make(PartGrid *g, int dim, const int cells[3], const int P[3], int rank, int replicated)
{
  g->dim = dim; g->rank = rank; g->replicated = replicated;
  g->nranks = P[0] * P[1] * P[2];
  int r = rank;
  for (int d = 0; d < 3; d++) {
    g->P[d] = d < dim ? P[d] : 1;
    g->coord[d] = r % P[d]; r /= P[d];
    g->nn[d] = d < dim ? cells[d] + 1 : 1;
    if (d < dim && cells[d] % g->P[d]) return 1;
    g->cpr[d] = d < dim ? cells[d] / g->P[d] : 1;
  }
  auto own_of = [&](const int c[3]) {
    PartBox b;
    for (int d = 0; d < 3; d++) {
      if (d < dim) part_own_range(c[d], g->cpr[d], &b.lo[d], &b.hi[d]);
      else { b.lo[d] = 0; b.hi[d] = 1; }
    }
    return b;
  };
  auto ext_of = [&](const PartBox &o) {
    PartBox e = o;
    for (int d = 0; d < dim; d++) { if (e.lo[d] > 0) e.lo[d]--; if (e.hi[d] < g->nn[d]) e.hi[d]++; }
    return e;
  };
  g->own = own_of(g->coord);
  g->ext = ext_of(g->own);
  for (int d = 0; d < 2; d++) {      // odd pitches in all but the slowest direction
    const int e = g->own.hi[d] - g->own.lo[d];
    g->pitch[d] = (!replicated && d < dim - 1 && e > 1 && (e % 2) == 0) ? e + 1 : e;
  }
  g->n_own = replicated ? g->nn[0] * g->nn[1] * g->nn[2]
                        : (box_count(g->own) == 0 ? 0 : g->pitch[0] * g->pitch[1] * (g->own.hi[2] - g->own.lo[2]));
  g->nnb = 0;
  g->nb_recv_off[0] = g->nb_send_off[0] = 0;
  for (int s = 0; s < 27; s++) g->nb_slot[s] = -1;
  if (!replicated) {
    // neighbours in ascending rank order (z slowest, x fastest = rank order)
    for (int dz = -1; dz <= 1; dz++) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
      if (!dx && !dy && !dz) continue;
      int c[3] = {g->coord[0] + dx, g->coord[1] + dy, g->coord[2] + dz};
      bool ok = true;
      for (int d = 0; d < 3; d++) if (c[d] < 0 || c[d] >= g->P[d]) ok = false;
      if (!ok) continue;
      PartBox o = own_of(c);
      PartBox recv = box_and(o, g->ext), send = box_and(g->own, ext_of(o));
      if (box_count(recv) == 0 && box_count(send) == 0) continue;
      int k = g->nnb++;
      g->nb_rank[k] = c[0] + g->P[0] * (c[1] + g->P[1] * c[2]);
      g->nb_slot[(dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)] = k;
      g->nb_recv[k] = recv; g->nb_send[k] = send;
      g->nb_recv_off[k + 1] = g->nb_recv_off[k] + box_count(recv);
      g->nb_send_off[k + 1] = g->nb_send_off[k] + box_count(send);
    }
  }
  g->n_ghost = g->nb_recv_off[g->nnb];
  return 0;
}

#endif
