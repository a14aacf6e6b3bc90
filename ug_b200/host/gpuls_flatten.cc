// gpuls_flatten.cc -- see gpuls_flatten.h.  Host-side product code (no CUDA, no oracle).
#include "config.h"
#include "gpuls_flatten.h"

#include "shapes.h"
#include "ugm.h"
#include "algebra.h"
#include "transgrid.h"      // CRITBIT

#include <cstdio>

USING_UG_NAMESPACES

namespace gpuls {

// The single vector type that carries the descriptor (P1/Q1 nodal formats: exactly one).
static int UniqueType(const VECDATA_DESC *vd)
{
  int t0 = -1;
  for (int t = 0; t < NVECTYPES; t++)
    if (VD_NCMPS_IN_TYPE(vd, t) > 0) {
      if (t0 >= 0) return -1;
      t0 = t;
    }
  return t0;
}

int NodeComps(const VECDATA_DESC *vd)
{
  int t0 = UniqueType(vd);
  if (t0 < 0) return -1;
  return VD_NCMPS_IN_TYPE(vd, t0);
}

// 0-based row numbers in list order.  Every Flatten* function numbers the grids it reads ITSELF: VINDEX is a scratch field that any host
// numproc may renumber between two calls (l_setindex of the reference is 1-based and is called by the host lu / gs / ilu PreProcess).
static int Renumber(GRID *g)
{
  int n = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v)) VINDEX(v) = n++;
  return n;
}

int FlattenFlags(MULTIGRID *mg, int level, const VECDATA_DESC *x, FlatLevel &out)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  int t0 = UniqueType(x);
  if (g == NULL || t0 < 0) return 1;
  out.bs = VD_NCMPS_IN_TYPE(x, t0);
  out.n = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v)) {
    if (VTYPE(v) != t0) return 2;  // mixed vector types are outside the hot-path scope
    VINDEX(v) = out.n++;           // l_setindex semantics (0-based here; ugiter.cc:169 starts at 1)
  }
  out.vclass.resize(out.n); out.vnclass.resize(out.n); out.ctl.resize(out.n); out.skip.resize(out.n);
  int r = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v), r++) {
    out.vclass[r]  = (uint8_t)VCLASS(v);
    out.vnclass[r] = (uint8_t)VNCLASS(v);
    out.ctl[r]     = (uint8_t)((NEW_DEFECT(v) ? 1u : 0u) | (FINE_GRID_DOF(v) ? 2u : 0u));
    out.skip[r]    = (uint32_t)VECSKIP(v);
  }
  return 0;
}

// Component index table of the (t0,t0) block, row-major: ugblas.cc:117-133 SET_MD_CMP_*.
static int BlockComps(const MATDATA_DESC *A, int t0, int bs, SHORT *comp)
{
  if (MD_ROWS_IN_RT_CT(A, t0, t0) != bs || MD_COLS_IN_RT_CT(A, t0, t0) != bs) return 1;
  for (int k = 0; k < bs * bs; k++) comp[k] = MD_MCMP_OF_RT_CT(A, t0, t0, k);
  return 0;
}

int FlattenMatrix(MULTIGRID *mg, int level, const MATDATA_DESC *A, FlatLevel &out)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  if (g == NULL || out.n <= 0) return 1;
  if (Renumber(g) != out.n) return 1;
  VECTOR *v0 = FIRSTVECTOR(g);
  int t0 = VTYPE(v0);
  int bs = out.bs, bb = bs * bs;
  SHORT comp[MAX_SINGLE_MAT_COMP];
  if (BlockComps(A, t0, bs, comp)) return 3;

  out.rowptr.assign(out.n + 1, 0);
  size_t nnz = 0;
  int r = 0;
  for (VECTOR *v = v0; v != NULL; v = SUCCVC(v), r++) {
    for (MATRIX *m = VSTART(v); m != NULL; m = MNEXT(m)) nnz++;
    out.rowptr[r + 1] = (int32_t)nnz;
  }
  out.col.resize(nnz);
  out.val.resize(nnz * bb);
  size_t e = 0;
  for (VECTOR *v = v0; v != NULL; v = SUCCVC(v)) {
    if (VSTART(v) != NULL && MDEST(VSTART(v)) != v) return 4;  // diag-first invariant (npcheck.cc CheckVector)
    for (MATRIX *m = VSTART(v); m != NULL; m = MNEXT(m), e++) {
      out.col[e] = (int32_t)VINDEX(MDEST(m));
      for (int k = 0; k < bb; k++) out.val[e * bb + k] = MVALUE(m, comp[k]);
    }
  }
  return 0;
}

int FlattenMatrixValues(MULTIGRID *mg, int level, const MATDATA_DESC *A, std::vector<double> &val, int bs)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  if (g == NULL) return 1;
  VECTOR *v0 = FIRSTVECTOR(g);
  int bb = bs * bs;
  SHORT comp[MAX_SINGLE_MAT_COMP];
  if (BlockComps(A, VTYPE(v0), bs, comp)) return 3;
  size_t e = 0;
  for (VECTOR *v = v0; v != NULL; v = SUCCVC(v))
    for (MATRIX *m = VSTART(v); m != NULL; m = MNEXT(m), e++) {
      if ((e + 1) * bb > val.size()) return 5;
      for (int k = 0; k < bb; k++) val[e * bb + k] = MVALUE(m, comp[k]);
    }
  return 0;
}

int ScatterMatrixValues(MULTIGRID *mg, int level, const MATDATA_DESC *A, const std::vector<double> &val, int bs)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  if (g == NULL) return 1;
  VECTOR *v0 = FIRSTVECTOR(g);
  int bb = bs * bs;
  SHORT comp[MAX_SINGLE_MAT_COMP];
  if (BlockComps(A, VTYPE(v0), bs, comp)) return 3;
  size_t e = 0;
  for (VECTOR *v = v0; v != NULL; v = SUCCVC(v))
    for (MATRIX *m = VSTART(v); m != NULL; m = MNEXT(m), e++) {
      if ((e + 1) * bb > val.size()) return 5;
      for (int k = 0; k < bb; k++) MVALUE(m, comp[k]) = val[e * bb + k];
    }
  return 0;
}

int ScatterSkip(MULTIGRID *mg, int level, const std::vector<uint32_t> &skip)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  if (g == NULL) return 1;
  size_t r = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v), r++) {
    if (r >= skip.size()) return 5;
    VECSKIP(v) = skip[r];
  }
  return 0;
}

int FlattenElements(MULTIGRID *mg, int level, std::vector<int64_t> &elem_ptr, std::vector<int32_t> &elem_row, std::vector<double> &coord,
                    std::vector<uint32_t> &dirichlet_skip, int bs)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  if (g == NULL) return 1;
  const int n = NVEC(g);
  elem_ptr.assign(1, 0); elem_row.clear();
  coord.assign((size_t)n * DIM, 0.0);
  dirichlet_skip.assign((size_t)n, 0u);
  for (NODE *nd = FIRSTNODE(g); nd != NULL; nd = SUCCN(nd)) {
    const int r = (int)VINDEX(NVECTOR(nd));
    if (r < 0 || r >= n) return 2;
    for (int d = 0; d < DIM; d++) coord[(size_t)r * DIM + d] = CVECT(MYVERTEX(nd))[d];
  }
  for (ELEMENT *e = FIRSTELEMENT(g); e != NULL; e = SUCCE(e)) {
    for (int i = 0; i < CORNERS_OF_ELEM(e); i++) {
      const int r = (int)VINDEX(NVECTOR(CORNER(e, i)));
      elem_row.push_back((int32_t)r);
      if (OBJT(e) == BEOBJ && OBJT(MYVERTEX(CORNER(e, i))) == BVOBJ) dirichlet_skip[r] = (1u << bs) - 1u;
    }
    elem_ptr.push_back((int64_t)elem_row.size());
  }
  return 0;
}

// Standard transfer stencils, mirroring the per-node logic of
// StandardIntCorNodeVector (transgrid.cc:269-307) for P and
// StandardRestrictNodeVector (transgrid.cc:150-189) for R.
int FlattenTransfer(MULTIGRID *mg, int level, FlatLevel &out)
{
  if (level <= 0) return 1;
  GRID *fg = GRID_ON_LEVEL(mg, level);
  GRID *cg = GRID_ON_LEVEL(mg, level - 1);
  if (fg == NULL || cg == NULL) return 1;
  int nf = out.n, nc = NVEC(cg);
  if (Renumber(fg) != nf || Renumber(cg) != nc) return 1;

  // P rows are indexed by vector row; build per-node first (one node per nodal vector).
  std::vector<int32_t> pcnt(nf, 0);
  struct Ent { int32_t c; double w; };
  std::vector<Ent> pent((size_t)nf * MAX_CORNERS_OF_ELEM);
  std::vector<int32_t> rcnt(nc + 1, 0);
  out.node_row.clear();

  DOUBLE c[MAX_CORNERS_OF_ELEM];
  for (NODE *nd = FIRSTNODE(fg); nd != NULL; nd = SUCCN(nd)) {
    VECTOR *v = NVECTOR(nd);
    int r = VINDEX(v);
    out.node_row.push_back(r);
    if (pcnt[r] != 0) return 6;  // periodic identification (several nodes per vector): out of scope
    Ent *pe = &pent[(size_t)r * MAX_CORNERS_OF_ELEM];
    int k = 0;
    if (CORNERTYPE(nd)) {
      pe[k].c = VINDEX(NVECTOR((NODE *)NFATHER(nd)));
      pe[k].w = 1.0;
      k++;
    } else {
      VERTEX *vx = MYVERTEX(nd);
      ELEMENT *el = VFATHER(vx);
      int n = CORNERS_OF_ELEM(el);
      GNs(n, LCVECT(vx), c);
      for (int i = 0; i < n; i++)
        if (c[i] != 0.0) {           // transgrid.cc:304
          pe[k].c = VINDEX(NVECTOR(CORNER(el, i)));
          pe[k].w = c[i];
          k++;
        }
    }
    pcnt[r] = k;
    if (VCLASS(v) >= NEWDEF_CLASS)   // transgrid.cc:153
      for (int i = 0; i < k; i++) rcnt[pe[i].c + 1]++;
  }

  out.p_rowptr.assign(nf + 1, 0);
  for (int r = 0; r < nf; r++) out.p_rowptr[r + 1] = out.p_rowptr[r] + pcnt[r];
  out.p_col.resize(out.p_rowptr[nf]);
  out.p_w.resize(out.p_rowptr[nf]);
  for (int r = 0; r < nf; r++) {
    const Ent *pe = &pent[(size_t)r * MAX_CORNERS_OF_ELEM];
    for (int i = 0; i < pcnt[r]; i++) {
      out.p_col[out.p_rowptr[r] + i] = pe[i].c;
      out.p_w[out.p_rowptr[r] + i] = pe[i].w;
    }
  }

  // R: coarse rows; contributions appended in fine NODE list order (the reference's scatter order).
  out.r_rowptr.assign(nc + 1, 0);
  for (int r = 0; r < nc; r++) out.r_rowptr[r + 1] = out.r_rowptr[r] + rcnt[r + 1];
  out.r_col.resize(out.r_rowptr[nc]);
  out.r_w.resize(out.r_rowptr[nc]);
  std::vector<int32_t> fill(out.r_rowptr.begin(), out.r_rowptr.end() - 1);
  for (NODE *nd = FIRSTNODE(fg); nd != NULL; nd = SUCCN(nd)) {
    VECTOR *v = NVECTOR(nd);
    if (VCLASS(v) < NEWDEF_CLASS) continue;
    int r = VINDEX(v);
    const Ent *pe = &pent[(size_t)r * MAX_CORNERS_OF_ELEM];
    for (int i = 0; i < pcnt[r]; i++) {
      int32_t pos = fill[pe[i].c]++;
      out.r_col[pos] = r;
      out.r_w[pos] = pe[i].w;
    }
  }
  return 0;
}

int FlattenTransferIMAT(MULTIGRID *mg, int level, FlatLevel &out)
{
  if (level <= BOTTOMLEVEL(mg)) return 1;          // levels < 1: the algebraic levels of an AMG transfer (np/procs/amgtransfer.cc), same lists
  GRID *fg = GRID_ON_LEVEL(mg, level), *cg = GRID_ON_LEVEL(mg, level - 1);
  if (fg == NULL || cg == NULL) return 2;
  const int nf = out.n, nc = NVEC(cg), bs = out.bs;
  if (Renumber(fg) != nf || Renumber(cg) != nc) return 2;
  out.p_rowptr.assign(nf + 1, 0);
  out.p_col.clear(); out.p_w.clear();
  out.node_row.clear();
  std::vector<int32_t> rcnt(nc + 1, 0);
  for (VECTOR *v = FIRSTVECTOR(fg); v != NULL; v = SUCCVC(v)) {
    const int r = VINDEX(v);
    for (int i = 0; i < bs; i++) if (CRITBIT(v, i)) return 8;
    for (MATRIX *m = VISTART(v); m != NULL; m = NEXT(m)) {
      const double c = MVALUE(m, 0);
      for (int i = 0; i < bs; i++)
        for (int j = 0; j < bs; j++)
          if (MVALUE(m, i * bs + j) != (i == j ? c : 0.0)) return 7;      // not c*I
      out.p_col.push_back(VINDEX(MDEST(m)));
      out.p_w.push_back(c);
      if (VCLASS(v) >= NEWDEF_CLASS) rcnt[VINDEX(MDEST(m)) + 1]++;        // transgrid.cc:1142
    }
    out.p_rowptr[r + 1] = (int32_t)out.p_col.size();
  }
  out.r_rowptr.assign(nc + 1, 0);
  for (int r = 0; r < nc; r++) out.r_rowptr[r + 1] = out.r_rowptr[r] + rcnt[r + 1];
  out.r_col.resize(out.r_rowptr[nc]);
  out.r_w.resize(out.r_rowptr[nc]);
  std::vector<int32_t> fill(out.r_rowptr.begin(), out.r_rowptr.end() - 1);
  for (VECTOR *v = FIRSTVECTOR(fg); v != NULL; v = SUCCVC(v)) {
    if (VCLASS(v) < NEWDEF_CLASS) continue;
    const int r = VINDEX(v);
    for (int e = out.p_rowptr[r]; e < out.p_rowptr[r + 1]; e++) {
      const int32_t pos = fill[out.p_col[e]]++;
      out.r_col[pos] = r;
      out.r_w[pos] = out.p_w[e];
    }
  }
  return 0;
}

void GatherVector(MULTIGRID *mg, int level, const VECDATA_DESC *vd, int bs, double *host)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  size_t k = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v)) {
    const SHORT *cmp = VD_CMPPTR_OF_TYPE(vd, VTYPE(v));
    for (int i = 0; i < bs; i++) host[k++] = VVALUE(v, cmp[i]);
  }
}

void ScatterVector(MULTIGRID *mg, int level, const VECDATA_DESC *vd, int bs, const double *host)
{
  GRID *g = GRID_ON_LEVEL(mg, level);
  size_t k = 0;
  for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v)) {
    const SHORT *cmp = VD_CMPPTR_OF_TYPE(vd, VTYPE(v));
    for (int i = 0; i < bs; i++) VVALUE(v, cmp[i]) = host[k++];
  }
}

}  // namespace gpuls
