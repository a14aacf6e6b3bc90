/* oracle/ugport.c -- TEST INFRASTRUCTURE ONLY: see ugport.h.
 *
 * Sequential, one statement per reference statement where it matters for rounding; compile with
 * -ffp-contract=off.  Each function cites the reference lines it restates (paths relative to the
 * reference root, UG 3.12.1).
 */
#include "ugport.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEWDEF_CLASS 2   /* gm/algebra.h:70 */
#define ACTIVE_CLASS 3   /* gm/algebra.h:71 */
#define CTL_NEW_DEFECT 1u
#define CTL_FINE_GRID_DOF 2u

static int row_on(const ugport_level *L, int rowmode, int r)
{
  if (rowmode == 0) return 1;
  if (rowmode == 1) return (L->ctl[r] & CTL_NEW_DEFECT) != 0;    /* vecloop.ct:31 */
  return (L->ctl[r] & CTL_FINE_GRID_DOF) != 0;                   /* vecloop.ct:24 */
}

/* np/algebra/matloop.ct:74-100 with T_MOD_SCAL ugblas.cc:4007 (scalar) and MATMUL_nn_SUCC
 * ugblas.h:161-282 (blocks): per entry e_i = (m_i0*y_0 + m_i1*y_1) + m_i2*y_2 ; s_i += e_i */
void ugport_dmatmul(const ugport_level *L, int op, int rowmode, double *x, const double *y)
{
  int bs = L->bs, bb = bs * bs;
  for (int r = 0; r < L->n; r++) {
    if (!row_on(L, rowmode, r)) continue;
    double s[UGPORT_MAX_BS] = {0.0, 0.0, 0.0};
    for (int e = L->rowptr[r]; e < L->rowptr[r + 1]; e++) {
      const double *m = L->val + (size_t)e * bb;
      const double *w = y + (size_t)L->col[e] * bs;
      for (int i = 0; i < bs; i++) {
        double acc = m[i * bs] * w[0];
        for (int j = 1; j < bs; j++) acc = acc + m[i * bs + j] * w[j];
        s[i] += acc;
      }
    }
    for (int i = 0; i < bs; i++) {
      double *xi = x + (size_t)r * bs + i;
      if (op == 0) { if (bs == 1) *xi = s[i]; else { *xi = 0.0; *xi += s[i]; } }  /* matmode.ct:76 T_CLEAR_X */
      else if (op == 1) *xi += s[i];
      else *xi -= s[i];
    }
  }
}

#define VLOOP(body) \
  for (int r = 0; r < L->n; r++) { if (!row_on(L, rowmode, r)) continue; \
    for (int i = 0; i < L->bs; i++) { size_t k = (size_t)r * L->bs + i; body; } }

void ugport_dset(const ugport_level *L, int rowmode, double *x, double a) { VLOOP(x[k] = a) }
void ugport_dcopy(const ugport_level *L, int rowmode, double *x, const double *y) { VLOOP(x[k] = y[k]) }
void ugport_dscal(const ugport_level *L, int rowmode, double *x, double a) { VLOOP(x[k] *= a) }
void ugport_dscalx(const ugport_level *L, int rowmode, double *x, const double *a) { VLOOP(x[k] *= a[i]) }
void ugport_dadd(const ugport_level *L, int rowmode, double *x, const double *y) { VLOOP(x[k] += y[k]) }
void ugport_dsub(const ugport_level *L, int rowmode, double *x, const double *y) { VLOOP(x[k] -= y[k]) }
void ugport_dminusadd(const ugport_level *L, int rowmode, double *x, const double *y) { VLOOP(x[k] = y[k] - x[k]) }
void ugport_daxpy(const ugport_level *L, int rowmode, double *x, double a, const double *y) { VLOOP(x[k] += a * y[k]) }
void ugport_daxpyx(const ugport_level *L, int rowmode, double *x, const double *a, const double *y) { VLOOP(x[k] += a[i] * y[k]) }
void ugport_ddot_acc(const ugport_level *L, int rowmode, const double *x, const double *y, double *sum) { VLOOP(*sum += x[k] * y[k]) }
void ugport_ddotx_acc(const ugport_level *L, int rowmode, const double *x, const double *y, double *sum) { VLOOP(sum[i] += x[k] * y[k]) }
void ugport_dnrm2_acc(const ugport_level *L, int rowmode, const double *x, double *sum) { VLOOP(double s = x[k]; *sum += s * s) }
void ugport_dnrm2x_acc(const ugport_level *L, int rowmode, const double *x, double *sum) { VLOOP(double s = x[k]; sum[i] += s * s) }

/* np/algebra/block.cc:104-142 SolveSmallBlock, n = 1,2,3 */
static int solve_small_block(int n, double *sol, const double *mat, const double *rhs)
{
  if (n == 1) { sol[0] = rhs[0] / mat[0]; return 0; }
  if (n == 2) {
    double det = mat[0] * mat[3] - mat[1] * mat[2];
    if (det == 0.0) return 1;
    det = 1.0 / det;
    sol[0] = (rhs[0] * mat[3] - rhs[1] * mat[1]) * det;
    sol[1] = (rhs[1] * mat[0] - rhs[0] * mat[2]) * det;
    return 0;
  }
  double M3div0 = mat[3] / mat[0];
  double M6div0 = mat[6] / mat[0];
  double aux = (mat[7] - M6div0 * mat[1]) / (mat[4] - M3div0 * mat[1]);
  sol[2] = (rhs[2] - M6div0 * rhs[0] - aux * (rhs[1] - M3div0 * rhs[0]))
           / (mat[8] - M6div0 * mat[2] - aux * (mat[5] - M3div0 * mat[2]));
  sol[1] = (rhs[1] - mat[3] / mat[0] * rhs[0] - (mat[5] - M3div0 * mat[2]) * sol[2])
           / (mat[4] - M3div0 * mat[1]);
  sol[0] = (rhs[0] - mat[1] * sol[1] - mat[2] * sol[2]) / mat[0];
  return 0;
}

/* np/algebra/ugiter.cc:271-335 */
int ugport_l_jac(const ugport_level *L, double *v, const double *d)
{
  int bs = L->bs, bb = bs * bs;
  for (int r = 0; r < L->n; r++) {
    double *vr = v + (size_t)r * bs;
    if (L->vclass[r] < ACTIVE_CLASS) { for (int i = 0; i < bs; i++) vr[i] = 0.0; continue; }
    double s[UGPORT_MAX_BS];
    for (int i = 0; i < bs; i++) s[i] = d[(size_t)r * bs + i];
    if (solve_small_block(bs, vr, L->val + (size_t)L->rowptr[r] * bb, s)) return 6; /* NUM_SMALL_DIAG */
  }
  return 0;
}

/* np/procs/iter.cc:817-842 Smoother with JacobiStep :911 */
int ugport_jac_smooth(const ugport_level *L, double *x, double *b, const double *damp)
{
  int err = ugport_l_jac(L, x, b);
  if (err) return err;
  ugport_dscalx(L, 0, x, damp);
  ugport_dmatmul(L, 2, 0, b, x);
  return 0;
}

/* np/algebra/ugiter.cc:412-615 (l_lgs), :735-930 (l_ugs), :1343-1560 (l_lsor), :1563-1775 (l_usor).  Row r = VINDEX.
 * Scalar descriptors (:441-470): sum = 0; sum += m*v over the off-diagonal entries in VSTART->MNEXT order whose column is on
 * the solved side and active; v = (d - sum)/diag, SOR: v = omega*(d - sum)/diag.  Block descriptors (:472-615): s = d;
 * s0..s2 = 0; MATMUL_nn per entry (ugblas.h:161-213: s_i += (m_i0*w_0 + m_i1*w_1) + m_i2*w_2); s -= s0..s2;
 * SolveSmallBlock; SOR: v_i *= omega_i (:1556). */
static int gs_sweep(const ugport_level *L, double *v, const double *d, int upper, const double *omega)
{
  int bs = L->bs, bb = bs * bs, n = L->n;
  for (int k = 0; k < n; k++) {
    int r = upper ? n - 1 - k : k;
    double *vr = v + (size_t)r * bs;
    if (L->vclass[r] < ACTIVE_CLASS) { for (int i = 0; i < bs; i++) vr[i] = 0.0; continue; }
    int e0 = L->rowptr[r];
    if (bs == 1) {
      double sum = 0.0;
      for (int e = e0 + 1; e < L->rowptr[r + 1]; e++) {
        int c = L->col[e];
        if ((upper ? c > r : c < r) && L->vclass[c] >= ACTIVE_CLASS) sum += L->val[e] * v[c];
      }
      if (omega) vr[0] = omega[0] * (d[r] - sum) / L->val[e0];
      else vr[0] = (d[r] - sum) / L->val[e0];
      continue;
    }
    double s[UGPORT_MAX_BS], acc[UGPORT_MAX_BS] = {0.0, 0.0, 0.0};
    for (int i = 0; i < bs; i++) s[i] = d[(size_t)r * bs + i];
    for (int e = e0 + 1; e < L->rowptr[r + 1]; e++) {
      int c = L->col[e];
      if (!((upper ? c > r : c < r) && L->vclass[c] >= ACTIVE_CLASS)) continue;
      const double *m = L->val + (size_t)e * bb, *w = v + (size_t)c * bs;
      for (int i = 0; i < bs; i++) {
        double t = m[i * bs] * w[0];
        for (int j = 1; j < bs; j++) t = t + m[i * bs + j] * w[j];
        acc[i] += t;
      }
    }
    for (int i = 0; i < bs; i++) s[i] -= acc[i];
    if (solve_small_block(bs, vr, L->val + (size_t)e0 * bb, s)) return 6; /* NUM_SMALL_DIAG */
    if (omega) for (int i = 0; i < bs; i++) vr[i] *= omega[i];
  }
  return 0;
}
int ugport_l_lgs(const ugport_level *L, double *v, const double *d) { return gs_sweep(L, v, d, 0, NULL); }
int ugport_l_ugs(const ugport_level *L, double *v, const double *d) { return gs_sweep(L, v, d, 1, NULL); }
int ugport_l_lsor(const ugport_level *L, double *v, const double *d, const double *omega) { return gs_sweep(L, v, d, 0, omega); }
int ugport_l_usor(const ugport_level *L, double *v, const double *d, const double *omega) { return gs_sweep(L, v, d, 1, omega); }

int ugport_smooth(const ugport_level *L, int kind, double *x, double *b, const double *damp, double *tmp)
{
  int err;
  switch (kind) {
  case UGPORT_SM_JAC: return ugport_jac_smooth(L, x, b, damp);
  case UGPORT_SM_GS:                                        /* Smoother iter.cc:817-842 with GSStep :1039 */
    if ((err = ugport_l_lgs(L, x, b))) return err;
    ugport_dscalx(L, 0, x, damp);
    ugport_dmatmul(L, 2, 0, b, x);
    return 0;
  case UGPORT_SM_SOR:                                       /* SORSmoother iter.cc:4786 (no dscalx: omega acts inside l_lsor) */
    if ((err = ugport_l_lsor(L, x, b, damp))) return err;
    ugport_dmatmul(L, 2, 0, b, x);
    return 0;
  case UGPORT_SM_ILU:                                       /* Smoother iter.cc:817-842 with ILUStep :5478 (L->ilu: ugport_ilu_decomp) */
    if (!L->ilu) return 1;
    if ((err = ugport_l_luiter(L, L->ilu, x, b))) return err;
    ugport_dscalx(L, 0, x, damp);
    ugport_dmatmul(L, 2, 0, b, x);
    return 0;
  case UGPORT_SM_SGS:                                       /* SGSSmoother iter.cc:1392-1456 */
    if ((err = ugport_l_lgs(L, tmp, b))) return err;
    ugport_dscalx(L, 0, tmp, damp);
    ugport_dmatmul(L, 2, 0, b, tmp);
    if ((err = ugport_l_ugs(L, x, b))) return err;
    ugport_dscalx(L, 0, x, damp);
    ugport_dmatmul(L, 2, 0, b, x);
    ugport_dadd(L, 0, x, tmp);
    return 0;
  }
  return 1;
}

/* np/algebra/transgrid.cc:117-189.  R rows list the contributions to one coarse vector in fine NODE
 * list order, so the serial sum below performs the same additions in the same order as the
 * reference's scatter loop. */
void ugport_restrict(const ugport_level *fine, const ugport_level *coarse, double *to, const double *from, const double *damp)
{
  int bs = fine->bs;
  for (int r = 0; r < coarse->n; r++) {
    double *tr = to + (size_t)r * bs;
    if (coarse->vnclass[r] >= NEWDEF_CLASS) for (int i = 0; i < bs; i++) tr[i] = 0.0;   /* :143-147 */
    uint32_t skip = coarse->skip[r];
    for (int e = fine->r_rowptr[r]; e < fine->r_rowptr[r + 1]; e++) {
      const double *f = from + (size_t)fine->r_col[e] * bs;
      double w = fine->r_w[e];
      for (int j = 0; j < bs; j++)
        if (!(skip & (1u << j))) {
          double s = damp[j] * f[j];            /* :164 / :173 */
          tr[j] += w * s;                       /* :165 (w = 1) / :187 */
        }
    }
  }
}

/* np/algebra/transgrid.cc:235-307 */
void ugport_interpolate(const ugport_level *fine, const ugport_level *coarse, double *to, const double *from, const double *damp)
{
  int bs = fine->bs;
  (void)coarse;
  for (int r = 0; r < fine->n; r++) {
    double *tr = to + (size_t)r * bs;
    for (int i = 0; i < bs; i++) tr[i] = 0.0;                                  /* :264-268 */
    uint32_t skip = fine->skip[r];
    for (int j = 0; j < bs; j++) {
      if (skip & (1u << j)) continue;
      for (int e = fine->p_rowptr[r]; e < fine->p_rowptr[r + 1]; e++)
        tr[j] += fine->p_w[e] * damp[j] * from[(size_t)fine->p_col[e] * bs + j];   /* :305 ; corner :290 */
    }
  }
}

/* IMAT mode.  RestrictByMatrix_General transgrid.cc:1113-1240: scalar descriptors :1127-1156 (w += m*v where VECSKIP(w) == 0, then
 * w *= damp[0] if damp[0] != 1), block descriptors :1158-1237 (per entry and component sum = 0; sum += m_ij*v_j; w_i += sum --
 * with the blocks c*I of the standard interpolation the sum is c*v_i; components with the skip bit are left out; finally
 * w_i *= damp_i on all rows if CheckDamp).  Both loops zero the coarse rows with VNCLASS >= NEWDEF_CLASS first and visit the fine
 * vectors with VCLASS >= NEWDEF_CLASS in list order: that is the order of the R rows in IMAT flattening. */
static int check_damp(int n, const double *damp) { for (int i = 0; i < n; i++) if (damp[i] != 1.0) return 1; return 0; }

void ugport_restrict_imat(const ugport_level *fine, const ugport_level *coarse, double *to, const double *from, const double *damp)
{
  int bs = fine->bs;
  for (int r = 0; r < coarse->n; r++) {
    double *tr = to + (size_t)r * bs;
    if (coarse->vnclass[r] >= NEWDEF_CLASS) for (int i = 0; i < bs; i++) tr[i] = 0.0;
    uint32_t skip = coarse->skip[r];
    for (int e = fine->r_rowptr[r]; e < fine->r_rowptr[r + 1]; e++) {
      const double *f = from + (size_t)fine->r_col[e] * bs;
      double w = fine->r_w[e];
      if (bs == 1) { if (skip == 0) tr[0] += w * f[0]; }
      else
        for (int i = 0; i < bs; i++)
          if (!(skip & (1u << i))) {
            double sum = 0.0;
            for (int j = 0; j < bs; j++) sum += (i == j ? w : 0.0) * f[j];
            tr[i] += sum;
          }
    }
  }
  if (check_damp(bs, damp))
    for (int r = 0; r < coarse->n; r++)
      if (coarse->vnclass[r] >= NEWDEF_CLASS) for (int i = 0; i < bs; i++) to[(size_t)r * bs + i] *= damp[i];
}

/* InterpolateCorrectionByMatrix_General transgrid.cc:1292-1390: to = 0; scalar :1306-1335 (rows with the skip bit stay 0);
 * blocks :1337-1382 (sum = 0; sum += m_ji*w_j; v_i += sum); then dscalx(to, damp) on ALL vectors if CheckDamp. */
void ugport_interpolate_imat(const ugport_level *fine, const ugport_level *coarse, double *to, const double *from, const double *damp)
{
  int bs = fine->bs;
  (void)coarse;
  for (int r = 0; r < fine->n; r++) {
    double *tr = to + (size_t)r * bs;
    for (int i = 0; i < bs; i++) tr[i] = 0.0;
    uint32_t skip = fine->skip[r];
    if (bs == 1 && (skip & 1u)) continue;
    for (int e = fine->p_rowptr[r]; e < fine->p_rowptr[r + 1]; e++) {
      const double *f = from + (size_t)fine->p_col[e] * bs;
      double w = fine->p_w[e];
      if (bs == 1) tr[0] += w * f[0];
      else
        for (int i = 0; i < bs; i++)
          if (!(skip & (1u << i))) {
            double sum = 0.0;
            for (int j = 0; j < bs; j++) sum += (i == j ? w : 0.0) * f[j];
            tr[i] += sum;
          }
    }
  }
  if (check_damp(bs, damp))
    for (int r = 0; r < fine->n; r++) for (int i = 0; i < bs; i++) to[(size_t)r * bs + i] *= damp[i];
}

/* ---- base level: `lu` (np/procs/iter.cc:6443 LUPreProcess, Smoother + LUStep) -------------------------------------------
 * l_lrdecomp (ugiter.cc:3657) eliminates on UG's per-row MATRIX LISTS and creates fill-in with CreateExtraConnection
 * (gm/algebra.cc:1101 -> CreateConnection :969), which inserts the two new MATRIX structs at the SECOND place of both row
 * lists (:1051-1078).  l_luiter (ugiter.cc:4444) then sums every row in list order, so the order of additions depends on the
 * order in which fill-in appeared -- and that depends on the numbers (a zero pivot, :3741 / :3829, or a zero correction block,
 * :3867, creates nothing).  The restatement therefore keeps the same ordered lists: row r = array of (column, block) in
 * VSTART->MNEXT order, new entries inserted at index 1. */
typedef struct { int col; double v[UGPORT_MAX_BS * UGPORT_MAX_BS]; } lu_ent;
typedef struct { int len, cap; lu_ent *e; } lu_row;
typedef struct { int n, bs; lu_row *row; } lu_fac;

static lu_ent *lu_find(lu_row *r, int c)           /* GetMatrix gm/algebra.cc */
{
  for (int k = 0; k < r->len; k++) if (r->e[k].col == c) return &r->e[k];
  return NULL;
}
static void lu_insert_second(lu_row *r, int c)     /* CreateConnection algebra.cc:1051-1078; new values are 0 */
{
  if (r->len == r->cap) { r->cap = r->cap ? 2 * r->cap : 8; r->e = (lu_ent *)realloc(r->e, sizeof(lu_ent) * (size_t)r->cap); }
  memmove(&r->e[2], &r->e[1], sizeof(lu_ent) * (size_t)(r->len - 1));
  memset(&r->e[1], 0, sizeof(lu_ent));
  r->e[1].col = c;
  r->len++;
}

/* InvertSmallBlock np/algebra/block.cc:272-321 (n = 1,2,3) */
static int invert_small_block(int n, const double *mat, double *inv)
{
  if (n == 1) { inv[0] = 1.0 / mat[0]; return 0; }
  if (n == 2) {
    double det = mat[0] * mat[3] - mat[1] * mat[2];
    if (det == 0.0) return 1;
    double invdet = 1.0 / det;
    inv[0] = mat[3] * invdet; inv[1] = -mat[1] * invdet; inv[2] = -mat[2] * invdet; inv[3] = mat[0] * invdet;
    return 0;
  }
  double det = mat[0] * mat[4] * mat[8] + mat[1] * mat[5] * mat[6] + mat[2] * mat[3] * mat[7]
               - mat[2] * mat[4] * mat[6] - mat[0] * mat[5] * mat[7] - mat[1] * mat[3] * mat[8];
  if (det == 0.0) return 1;
  double invdet = 1.0 / det;
  inv[0] = ( mat[4] * mat[8] - mat[5] * mat[7]) * invdet;
  inv[3] = (-mat[3] * mat[8] + mat[5] * mat[6]) * invdet;
  inv[6] = ( mat[3] * mat[7] - mat[4] * mat[6]) * invdet;
  inv[1] = (-mat[1] * mat[8] + mat[2] * mat[7]) * invdet;
  inv[4] = ( mat[0] * mat[8] - mat[2] * mat[6]) * invdet;
  inv[7] = (-mat[0] * mat[7] + mat[1] * mat[6]) * invdet;
  inv[2] = ( mat[1] * mat[5] - mat[2] * mat[4]) * invdet;
  inv[5] = (-mat[0] * mat[5] + mat[2] * mat[3]) * invdet;
  inv[8] = ( mat[0] * mat[4] - mat[1] * mat[3]) * invdet;
  return 0;
}

/* ---- ILU (SURVEY.md 8f.2 rest) ---------------------------------------------------------------------------------------------
 * l_ilubthdecomp np/algebra/ugiter.cc:2252-2640 as class `ilu` calls it (ILUPreProcess np/procs/iter.cc:5468: beta = the class's
 * $beta, threshold = rest = oldrestthresh = NULL, so VCUSED = 0 and no connection is ever created): incomplete decomposition on
 * the pattern of A, in place on `val` (a copy of the level's values in the same entry order), right-looking like the reference:
 * for every active row i in index order the diagonal (block) is inverted and STORED inverted (StoreInverse, :139), every entry
 * (j,i), j > i active, becomes the pivot M_ji * D_i^-1, and row j receives -pivot * M_ik for the active k > i of row i IN ROW i's
 * LIST ORDER -- on the pattern as a subtraction, off the pattern as the beta-modification of D_j (scalar :2418-2420: D_j +=
 * beta*|pivot*M_ik| entry by entry; blocks :2598-2640: row sums of |Cor| per (i,j) pair, then D_j columns scaled by
 * 1 + beta_l*RowSum_l).  Returns 0, or -(i+1) when the diagonal of row i cannot be inverted (the reference returns -i). */
static int csr_find(const ugport_level *L, int r, int c)          /* GetMatrix gm/algebra.cc: the diagonal is the row's first entry */
{
  for (int e = L->rowptr[r]; e < L->rowptr[r + 1]; e++) if (L->col[e] == c) return e;
  return -1;
}

int ugport_ilu_decomp(const ugport_level *L, const double *beta, double *val)
{
  int bs = L->bs, bb = bs * bs, n = L->n;
  memcpy(val, L->val, sizeof(double) * (size_t)L->rowptr[n] * bb);          /* dmatcopy(L, A) iter.cc:5461 */
  for (int i = 0; i < n; i++) {
    if (L->vclass[i] < ACTIVE_CLASS) continue;
    int e0 = L->rowptr[i], e1 = L->rowptr[i + 1];
    if (bs == 1) {
      double diag = val[e0];
      if (fabs(diag) < 2.220446049250313e-16 * 10 * 1e-20) return -(i + 1);   /* SMALL_D*1e-20, low/architecture.h:27-34 */
      double invdiag = 1.0 / diag;
      val[e0] = invdiag;
      for (int e = e0 + 1; e < e1; e++) {
        int j = L->col[e];
        if (!(L->vclass[j] >= ACTIVE_CLASS && j > i)) continue;
        int ji = csr_find(L, j, i);                                            /* MADJ(Mij) */
        if (ji < 0) return 1;
        double pivot = val[ji] * invdiag;
        val[ji] = pivot;
        if (pivot == 0.0) continue;
        for (int f = e0 + 1; f < e1; f++) {
          int k = L->col[f];
          if (!(L->vclass[k] >= ACTIVE_CLASS && k > i)) continue;
          int jk = csr_find(L, j, k);
          if (jk >= 0) val[jk] -= pivot * val[f];
          else if (beta) val[L->rowptr[j]] += beta[0] * fabs(pivot * val[f]);
        }
      }
      continue;
    }
    double inv[9], piv[9], cor[9], rowsum[UGPORT_MAX_BS], dampf[UGPORT_MAX_BS];
    double *Diag = val + (size_t)e0 * bb;
    if (invert_small_block(bs, Diag, inv)) return -(i + 1);
    for (int l = 0; l < bb; l++) Diag[l] = inv[l];
    for (int e = e0 + 1; e < e1; e++) {
      int j = L->col[e];
      if (!(L->vclass[j] >= ACTIVE_CLASS && j > i)) continue;
      for (int l = 0; l < bs; l++) rowsum[l] = 0.0;
      int ji = csr_find(L, j, i);
      if (ji < 0) return 1;
      double *Piv = val + (size_t)ji * bb;
      int pivzero = 1;
      for (int i0 = 0; i0 < bs; i0++)
        for (int j0 = 0; j0 < bs; j0++) {
          double sum = 0.0;
          for (int k0 = 0; k0 < bs; k0++) sum += Piv[i0 * bs + k0] * inv[k0 * bs + j0];
          piv[i0 * bs + j0] = sum;
          if (sum != 0.0) pivzero = 0;
        }
      for (int l = 0; l < bb; l++) Piv[l] = piv[l];
      if (pivzero) continue;
      for (int f = e0 + 1; f < e1; f++) {
        int k = L->col[f];
        if (!(L->vclass[k] >= ACTIVE_CLASS && k > i)) continue;
        const double *Elm = val + (size_t)f * bb;
        int corzero = 1;
        for (int i0 = 0; i0 < bs; i0++)
          for (int j0 = 0; j0 < bs; j0++) {
            double sum = 0.0;
            for (int k0 = 0; k0 < bs; k0++) sum += piv[i0 * bs + k0] * Elm[k0 * bs + j0];
            cor[i0 * bs + j0] = sum;
            if (sum != 0.0) corzero = 0;
          }
        if (corzero) continue;
        int jk = csr_find(L, j, k);
        if (jk >= 0) {
          double *Mat = val + (size_t)jk * bb;
          for (int l = 0; l < bb; l++) Mat[l] -= cor[l];
        } else {
          for (int l = 0; l < bs; l++)
            for (int m = 0; m < bs; m++) rowsum[l] += fabs(cor[l * bs + m] * 1.0 * 1.0);   /* j_/k_Normalization = 1 without VD_rest */
        }
      }
      if (!beta) continue;
      for (int m = 0; m < bs; m++) dampf[m] = 1.0 + beta[m] * rowsum[m];
      double *Djj = val + (size_t)L->rowptr[j] * bb;
      for (int m = 0; m < bs; m++)
        for (int l = 0; l < bs; l++) Djj[m * bs + l] *= dampf[l];
    }
  }
  return 0;
}

/* l_luiter ugiter.cc:4444-4796 on the decomposed values `lval` (pattern of the level): L v = d with Diag(L) = I over the active
 * columns c < r in list order (inactive rows get 0), then U v = v over the active c > r, the stored inverse diagonal applied by
 * multiplication (scalar :4510; blocks SolveInverseSmallBlock block.cc:225-253). */
int ugport_l_luiter(const ugport_level *L, const double *lval, double *v, const double *d)
{
  int bs = L->bs, bb = bs * bs, n = L->n;
  for (int pass = 0; pass < 2; pass++)
    for (int q = 0; q < n; q++) {
      int r = pass ? n - 1 - q : q;
      double *vr = v + (size_t)r * bs;
      if (L->vclass[r] < ACTIVE_CLASS) { if (!pass) for (int i = 0; i < bs; i++) vr[i] = 0.0; continue; }
      double acc[UGPORT_MAX_BS] = {0.0, 0.0, 0.0}, s[UGPORT_MAX_BS];
      for (int e = L->rowptr[r] + 1; e < L->rowptr[r + 1]; e++) {
        int c = L->col[e];
        if (!((pass ? c > r : c < r) && L->vclass[c] >= ACTIVE_CLASS)) continue;
        const double *m = lval + (size_t)e * bb, *w = v + (size_t)c * bs;
        for (int i = 0; i < bs; i++) {
          double t = m[i * bs] * w[0];
          for (int j = 1; j < bs; j++) t = t + m[i * bs + j] * w[j];
          acc[i] += t;
        }
      }
      if (!pass) { for (int i = 0; i < bs; i++) vr[i] = d[(size_t)r * bs + i] - acc[i]; continue; }
      for (int i = 0; i < bs; i++) s[i] = vr[i] - acc[i];
      const double *inv = lval + (size_t)L->rowptr[r] * bb;
      if (bs == 1) vr[0] = s[0] * inv[0];
      else
        for (int i = 0; i < bs; i++) {
          double sum = 0.0;
          for (int j = 0; j < bs; j++) sum += inv[i * bs + j] * s[j];
          vr[i] = sum;
        }
    }
  return 0;
}

/* ---- Galerkin coarse-grid operator ---------------------------------------------------------------------------
 * AssembleGalerkinByMatrix np/algebra/transgrid.cc:1575-1700 with symmetric = 0, after dmatset(coarse, 0) as `npcheck $G` calls it
 * (npcheck.cc:375-379): coarse += P^T A P on the stored interpolation matrices, accumulated in the reference's traversal order --
 * fine rows v in list order, their entries m = (v,w) in list order, the interpolation entries im of v in VISTART order, those jm
 * of w in VISTART order: coarse(iv,jv) += (m * im) * jm (scalar :1601-1627); blocks :1629-1700: sum over k,l of
 * (IM[i][k] * M[k][l]) * JM[j][l] with the c*I interpolation blocks written out (their zeros take part in the sums), then
 * coarse(iv,jv)[i][j] += sum.  `fine` / `coarse`: patterns, P of the fine level in IMAT order (dumps written with --imat);
 * fine_val: the fine matrix' values; coarse_val[nnz_c*bs*bs] receives the result.  The reference creates connections the coarse
 * pattern lacks (CreateExtraConnection); on nested geometric hierarchies the product stays on the pattern, and this restatement
 * returns 2 instead of growing it (pattern growth unpinned: no dump exercises it). */
int ugport_galerkin(const ugport_level *fine, const ugport_level *coarse, const double *fine_val, double *coarse_val)
{
  int bs = fine->bs, bb = bs * bs;
  memset(coarse_val, 0, sizeof(double) * (size_t)coarse->rowptr[coarse->n] * bb);       /* dmatset(level-1, A, 0.0) */
  for (int v = 0; v < fine->n; v++)
    for (int e = fine->rowptr[v]; e < fine->rowptr[v + 1]; e++) {
      int w = fine->col[e];
      const double *M = fine_val + (size_t)e * bb;
      for (int ie = fine->p_rowptr[v]; ie < fine->p_rowptr[v + 1]; ie++) {
        int iv = fine->p_col[ie];
        for (int je = fine->p_rowptr[w]; je < fine->p_rowptr[w + 1]; je++) {
          int jv = fine->p_col[je];
          int cm = csr_find(coarse, iv, jv);
          if (cm < 0) return 2;
          if (bs == 1) {
            double fac = M[0] * fine->p_w[ie];                                        /* :1609, hoisted out of the jm loop there too */
            coarse_val[cm] += fac * fine->p_w[je];
            continue;
          }
          double IM[9], JM[9];
          for (int a = 0; a < bb; a++) { IM[a] = 0.0; JM[a] = 0.0; }
          for (int a = 0; a < bs; a++) { IM[a * bs + a] = fine->p_w[ie]; JM[a * bs + a] = fine->p_w[je]; }
          double *C = coarse_val + (size_t)cm * bb;
          for (int i = 0; i < bs; i++)
            for (int j = 0; j < bs; j++) {
              double sum = 0.0;
              for (int k = 0; k < bs; k++)
                for (int l = 0; l < bs; l++) sum += IM[i * bs + k] * M[k * bs + l] * JM[j * bs + l];
              C[i * bs + j] += sum;
            }
        }
      }
    }
  return 0;
}

/* The pattern AssembleGalerkinByMatrix leaves on the coarse level when connections are missing (transgrid.cc:1615-1617, :1649-1651):
 * CreateExtraConnection -> CreateConnection (gm/algebra.cc:969-1081) puts BOTH matrices of a new connection at the SECOND place of their
 * rows' lists (:1051-1078), the diagonal stays first.  A row therefore ends up as: diagonal, the connections created by the product in
 * REVERSE order of creation, then the off-diagonal entries it had before.  Creation order = order of the first term (iv, jv) or (jv, iv)
 * in the traversal of ugport_galerkin.  start_rowptr / start_col: the pattern before the product (NULL: one diagonal entry per row, what
 * the AMG's GenerateNewGrid creates, np/algebra/amgtools.c:538-640).  Call with out_col == NULL to get the row pointers (sizes) first.
 * Plain dynamic rows and linear searches: test infrastructure for small levels. */
int ugport_galerkin_pattern(int nf, int nc, const int32_t *a_rowptr, const int32_t *a_col, const int32_t *p_rowptr, const int32_t *p_col,
                            const int32_t *start_rowptr, const int32_t *start_col, int32_t *out_rowptr, int32_t *out_col)
{
  typedef struct { int32_t *e; int n0, nnew, cap; } prow;    /* e[0..n0): the old entries, e[n0..n0+nnew): new ones in creation order */
  prow *R = (prow *)calloc((size_t)(nc > 0 ? nc : 1), sizeof(prow));
  for (int i = 0; i < nc; i++) {
    int n0 = start_rowptr ? start_rowptr[i + 1] - start_rowptr[i] : 1;
    R[i].cap = n0 + 8; R[i].n0 = n0; R[i].nnew = 0;
    R[i].e = (int32_t *)malloc(sizeof(int32_t) * (size_t)R[i].cap);
    if (start_rowptr) memcpy(R[i].e, start_col + start_rowptr[i], sizeof(int32_t) * (size_t)n0); else R[i].e[0] = i;
    if (n0 < 1 || R[i].e[0] != i) { for (int k = 0; k <= i; k++) free(R[k].e); free(R); return 4; }      /* diagonal first */
  }
#define PUSH(row, colv) do { prow *q = &R[row]; if (q->n0 + q->nnew == q->cap) { q->cap *= 2; q->e = (int32_t *)realloc(q->e, sizeof(int32_t) * (size_t)q->cap); } \
                             q->e[q->n0 + q->nnew++] = (colv); } while (0)
  for (int v = 0; v < nf; v++)
    for (int e = a_rowptr[v]; e < a_rowptr[v + 1]; e++) {
      int w = a_col[e];
      for (int ie = p_rowptr[v]; ie < p_rowptr[v + 1]; ie++) {
        int iv = p_col[ie];
        for (int je = p_rowptr[w]; je < p_rowptr[w + 1]; je++) {
          int jv = p_col[je], have = 0;
          const prow *q = &R[iv];
          for (int k = 0; k < q->n0 + q->nnew; k++) if (q->e[k] == jv) { have = 1; break; }      /* GetMatrix(iv, jv) */
          if (have) continue;
          PUSH(iv, jv); PUSH(jv, iv);                                                           /* one CONNECTION = both directions */
        }
      }
    }
#undef PUSH
  out_rowptr[0] = 0;
  for (int i = 0; i < nc; i++) out_rowptr[i + 1] = out_rowptr[i] + R[i].n0 + R[i].nnew;
  if (out_col)
    for (int i = 0; i < nc; i++) {
      int32_t *o = out_col + out_rowptr[i];
      *o++ = R[i].e[0];
      for (int k = R[i].nnew - 1; k >= 0; k--) *o++ = R[i].e[R[i].n0 + k];
      for (int k = 1; k < R[i].n0; k++) *o++ = R[i].e[k];
    }
  for (int i = 0; i < nc; i++) free(R[i].e);
  free(R);
  return 0;
}

void ugport_base_free(double *lu)
{
  lu_fac *F = (lu_fac *)lu;
  if (!F) return;
  for (int r = 0; r < F->n; r++) free(F->row[r].e);
  free(F->row); free(F);
}

/* l_lrdecomp ugiter.cc:3657-3880: scalar descriptors :3715-3768, block descriptors :3771-3880.  The returned handle is opaque. */
double *ugport_base_factor(const ugport_level *L)
{
  int n = L->n, bs = L->bs, bb = bs * bs;
  lu_fac *F = (lu_fac *)calloc(1, sizeof(lu_fac));
  F->n = n; F->bs = bs;
  F->row = (lu_row *)calloc((size_t)(n > 0 ? n : 1), sizeof(lu_row));
  for (int r = 0; r < n; r++) {                 /* dmatcopy(L, A): same lists */
    lu_row *R = &F->row[r];
    R->len = R->cap = L->rowptr[r + 1] - L->rowptr[r];
    R->e = (lu_ent *)calloc((size_t)(R->cap > 0 ? R->cap : 1), sizeof(lu_ent));
    for (int k = 0; k < R->len; k++) {
      int e = L->rowptr[r] + k;
      R->e[k].col = L->col[e];
      memcpy(R->e[k].v, L->val + (size_t)e * bb, sizeof(double) * (size_t)bb);
    }
  }
#define ACTIVE(x) (L->vclass[x] >= ACTIVE_CLASS)
  for (int i = 0; i < n; i++) {
    if (!ACTIVE(i)) continue;
    lu_row *Ri = &F->row[i];
    double inv[UGPORT_MAX_BS * UGPORT_MAX_BS];
    if (bs == 1) inv[0] = 1.0 / Ri->e[0].v[0];
    else if (invert_small_block(bs, Ri->e[0].v, inv)) { ugport_base_free((double *)F); return NULL; }
    memcpy(Ri->e[0].v, inv, sizeof(double) * (size_t)bb);                     /* StoreInverse */
    for (int a = 1; a < Ri->len; a++) {                                        /* Mij */
      int j = Ri->e[a].col;
      if (!(ACTIVE(j) && j > i)) continue;
      lu_ent *Mji = lu_find(&F->row[j], i);                                    /* MADJ(Mij) */
      double piv[UGPORT_MAX_BS * UGPORT_MAX_BS];
      int piv_zero = 1;
      if (bs == 1) { piv[0] = Mji->v[0] * inv[0]; piv_zero = piv[0] == 0.0; }
      else
        for (int i0 = 0; i0 < bs; i0++) for (int j0 = 0; j0 < bs; j0++) {
          double sum = 0.0;
          for (int k0 = 0; k0 < bs; k0++) sum += Mji->v[i0 * bs + k0] * inv[k0 * bs + j0];
          piv[i0 * bs + j0] = sum;
          if (sum != 0.0) piv_zero = 0;
        }
      memcpy(Mji->v, piv, sizeof(double) * (size_t)bb);
      if (piv_zero) continue;
      for (int c = 1; c < Ri->len; c++) {                                      /* Mik */
        int k = Ri->e[c].col;
        if (!(ACTIVE(k) && k > i)) continue;
        double cor[UGPORT_MAX_BS * UGPORT_MAX_BS];
        if (bs == 1) cor[0] = piv[0] * Ri->e[c].v[0];
        else {
          int cor_zero = 1;
          for (int i0 = 0; i0 < bs; i0++) for (int j0 = 0; j0 < bs; j0++) {
            double sum = 0.0;
            for (int k0 = 0; k0 < bs; k0++) sum += piv[i0 * bs + k0] * Ri->e[c].v[k0 * bs + j0];
            cor[i0 * bs + j0] = sum;
            if (sum != 0.0) cor_zero = 0;
          }
          if (cor_zero) continue;
        }
        lu_ent *Mjk = lu_find(&F->row[j], k);
        if (!Mjk) {                                                            /* CreateExtraConnection(g, vj, vk) */
          lu_insert_second(&F->row[j], k);
          lu_insert_second(&F->row[k], j);
          Mjk = &F->row[j].e[1];
        }
        for (int q = 0; q < bb; q++) Mjk->v[q] -= cor[q];
      }
    }
  }
#undef ACTIVE
  return (double *)F;
}

/* l_luiter ugiter.cc:4444-4795 on the factored lists: scalar descriptors :4470-4518, block descriptors :4522-4795 */
static void base_lu_solve(const ugport_level *L, const double *lu, double *v, const double *d)
{
  const lu_fac *F = (const lu_fac *)lu;
  int n = F->n, bs = F->bs;
  for (int r = 0; r < n; r++) {
    double *vr = v + (size_t)r * bs;
    if (L->vclass[r] < ACTIVE_CLASS) { for (int i = 0; i < bs; i++) vr[i] = 0.0; continue; }
    const lu_row *R = &F->row[r];
    double acc[UGPORT_MAX_BS] = {0.0, 0.0, 0.0};
    for (int a = 1; a < R->len; a++) {
      int c = R->e[a].col;
      if (!(c < r && L->vclass[c] >= ACTIVE_CLASS)) continue;
      const double *m = R->e[a].v, *w = v + (size_t)c * bs;
      for (int i = 0; i < bs; i++) {
        double t = m[i * bs] * w[0];
        for (int j = 1; j < bs; j++) t = t + m[i * bs + j] * w[j];
        acc[i] += t;
      }
    }
    for (int i = 0; i < bs; i++) vr[i] = d[(size_t)r * bs + i] - acc[i];          /* Diag(L) = I */
  }
  for (int r = n - 1; r >= 0; r--) {
    if (L->vclass[r] < ACTIVE_CLASS) continue;
    double *vr = v + (size_t)r * bs;
    const lu_row *R = &F->row[r];
    double acc[UGPORT_MAX_BS] = {0.0, 0.0, 0.0}, s[UGPORT_MAX_BS];
    for (int a = 1; a < R->len; a++) {
      int c = R->e[a].col;
      if (!(c > r && L->vclass[c] >= ACTIVE_CLASS)) continue;
      const double *m = R->e[a].v, *w = v + (size_t)c * bs;
      for (int i = 0; i < bs; i++) {
        double t = m[i * bs] * w[0];
        for (int j = 1; j < bs; j++) t = t + m[i * bs + j] * w[j];
        acc[i] += t;
      }
    }
    for (int i = 0; i < bs; i++) s[i] = vr[i] - acc[i];
    const double *inv = R->e[0].v;
    if (bs == 1) vr[0] = s[0] * inv[0];                                             /* :4510 */
    else
      for (int i = 0; i < bs; i++) {                                                /* SolveInverseSmallBlock block.cc:225-253 */
        double sum = 0.0;
        for (int j = 0; j < bs; j++) sum += inv[i * bs + j] * s[j];
        vr[i] = sum;
      }
  }
}

static int sc_cmp(const double *x, const double *y, int n)   /* npscan.cc:1027 */
{
  for (int i = 0; i < n; i++) if (fabs(x[i]) >= fabs(y[i])) return 0;
  return 1;
}

/* base `ls $I lu`: LinearResiduum + LinearSolver (ls.cc:577,637) with Iter = Smoother/ILUStep (damp 1) */
static void base_solve(const ugport_level *L, const ugport_cfg *cfg, const double *lu, double *c, double *b, double *t)
{
  int bs = L->bs, N = L->n * bs;
  double first[UGPORT_MAX_BS], last[UGPORT_MAX_BS], reach[UGPORT_MAX_BS], absl[UGPORT_MAX_BS];
  memset(last, 0, sizeof last);
  ugport_dnrm2x_acc(L, 1, b, last);           /* Residuum at tl = base level: NEW_DEFECT rows (vecloop.ct:30-38) */
  for (int i = 0; i < bs; i++) { last[i] = sqrt(last[i]); first[i] = last[i]; absl[i] = cfg->base_abslimit;
    reach[i] = first[i] * cfg->base_reduction; if (reach[i] == 0.0) reach[i] = cfg->base_reduction; }
  if (sc_cmp(first, absl, bs)) return;
  double *cc = (double *)malloc(sizeof(double) * N);
  for (int it = 0; it < cfg->base_maxit; it++) {
    (void)t;
    base_lu_solve(L, lu, cc, b);                 /* dset(c,0); Iter: c' = LU^-1 b; dscalx(1); b -= A c' */
    ugport_dmatmul(L, 2, 0, b, cc);
    for (int k = 0; k < N; k++) c[k] += cc[k];   /* LSUpdate ls.cc:869 */
    memset(last, 0, sizeof last);
    ugport_dnrm2x_acc(L, 1, b, last);
    for (int i = 0; i < bs; i++) last[i] = sqrt(last[i]);
    if (sc_cmp(last, absl, bs) || sc_cmp(last, reach, bs)) break;
  }
  free(cc);
}

/* np/procs/transfer.cc:488-516 */
void ugport_minimize_level(const ugport_level *L, double *c, double *b, double *t)
{
  double a0 = 0.0, a1 = 0.0;
  ugport_dmatmul(L, 0, 0, t, c);                 /* :498 dmatmul(t, A, c) */
  ugport_ddot_acc(L, 0, t, b, &a0);              /* :504 */
  ugport_dnrm2_acc(L, 0, t, &a1);                /* :506 dnrm2 = sqrt of the sum ... */
  a1 = sqrt(a1);
  a1 *= a1;                                      /* :508 ... squared again */
  ugport_dscal(L, 0, c, 1 + a0 / a1);            /* :511 */
  ugport_daxpy(L, 0, b, -a0 / a1, t);            /* :513 */
}

/* np/procs/iter.cc:7741-7949 */
int ugport_lmgc(const ugport_level *lv, const ugport_cfg *cfg, const double *lu, int level, double **c, double **b, double **t)
{
  const ugport_level *L = &lv[level];
  double one[UGPORT_MAX_BS] = {1.0, 1.0, 1.0};
  if (level <= cfg->baselevel) {
    if (cfg->base_hook) return cfg->base_hook(cfg->base_user, level, c[level], b[level]);
    base_solve(L, cfg, lu, c[level], b[level], t[level]);
    return 0;
  }
  double *tmp = cfg->smoother == UGPORT_SM_SGS ? (double *)malloc(sizeof(double) * (size_t)L->n * L->bs) : NULL;
  for (int i = 0; i < cfg->nu1; i++) {
    int err = ugport_smooth(L, cfg->smoother, t[level], b[level], cfg->smooth_damp, tmp);
    if (err) { free(tmp); return err; }
    ugport_dadd(L, 0, c[level], t[level]);
  }
  const int imat = cfg->imat || level <= cfg->imat_below;
  if (imat) ugport_restrict_imat(L, &lv[level - 1], b[level - 1], b[level], one);
  else ugport_restrict(L, &lv[level - 1], b[level - 1], b[level], one);           /* :7843, Factor_One */
  ugport_dset(&lv[level - 1], 0, c[level - 1], 0.0);                         /* :7873 */
  for (int g = 0; g < cfg->gamma; g++) {
    int err = ugport_lmgc(lv, cfg, lu, level - 1, c, b, t);
    if (err) { free(tmp); return err; }
  }
  if (imat) ugport_interpolate_imat(L, &lv[level - 1], t[level], c[level - 1], cfg->cycle_damp);
  else ugport_interpolate(L, &lv[level - 1], t[level], c[level - 1], cfg->cycle_damp);  /* :7886 */
  ugport_dadd(L, 0, c[level], t[level]);                                     /* :7903 */
  ugport_dmatmul(L, 2, 0, b[level], t[level]);                               /* :7905 */
  for (int i = 0; i < cfg->nu2; i++) {
    int err = ugport_smooth(L, cfg->smoother, t[level], b[level], cfg->smooth_damp, tmp);
    if (err) { free(tmp); return err; }
    ugport_dadd(L, 0, c[level], t[level]);
  }
  free(tmp);
  if (cfg->level_opt) ugport_minimize_level(L, c[level], b[level], t[level]);   /* :7944 AdaptCorrection -> transfer.cc:812 */
  return 0;
}

/* ls.cc:562: dmatmul_minus(bl..level, ON_SURFACE); surface loops: matloop.ct:22-73 */
void ugport_ls_defect(const ugport_level *lv, int fr, int bl, int level, double **x, double **b)
{
  (void)bl;
  for (int l = fr; l < level; l++) ugport_dmatmul(&lv[l], 2, 2, b[l], x[l]);
  ugport_dmatmul(&lv[level], 2, 1, b[level], x[level]);
}

/* ls.cc:577: dnrm2x(bl..level, ON_SURFACE) -- one running sum across levels, then sqrt */
void ugport_ls_residuum(const ugport_level *lv, int fr, int bl, int level, double **b, double *defect)
{
  (void)bl;
  int bs = lv[level].bs;
  for (int i = 0; i < bs; i++) defect[i] = 0.0;
  for (int l = fr; l < level; l++) ugport_dnrm2x_acc(&lv[l], 2, b[l], defect);
  ugport_dnrm2x_acc(&lv[level], 1, b[level], defect);
  for (int i = 0; i < bs; i++) defect[i] = sqrt(defect[i]);
}

/* ls.cc:637-749 with Update = LSUpdate (:869, dadd on levels baselevel..level) */
int ugport_solve(const ugport_level *lv, const ugport_cfg *cfg, int fr, int level, double **x, double **b, double **c, double **t,
                 int maxiter, const double *abslimit, const double *reduction, double *first_defect, double *history)
{
  int bs = lv[level].bs, bl = cfg->baselevel, it, done = 0;
  double last[UGPORT_MAX_BS], reach[UGPORT_MAX_BS];
  double *lu = cfg->base_hook ? NULL : ugport_base_factor(&lv[bl]);
  ugport_ls_residuum(lv, fr, bl, level, b, last);
  for (int i = 0; i < bs; i++) { first_defect[i] = last[i]; reach[i] = last[i] * reduction[i]; if (reach[i] == 0.0) reach[i] = reduction[i]; }
  if (sc_cmp(last, abslimit, bs)) { ugport_base_free(lu); return 0; }
  for (it = 0; it < maxiter; it++) {
    ugport_dset(&lv[level], 0, c[level], 0.0);
    if (ugport_lmgc(lv, cfg, lu, level, c, b, t)) { ugport_base_free(lu); return -1; }
    for (int l = bl; l <= level; l++) ugport_dadd(&lv[l], 0, x[l], c[l]);
    ugport_ls_residuum(lv, fr, bl, level, b, last);
    if (history) for (int i = 0; i < bs; i++) history[it * bs + i] = last[i];
    done = it + 1;
    if (sc_cmp(last, abslimit, bs) || sc_cmp(last, reach, bs)) break;
  }
  ugport_base_free(lu);
  return done;
}

/* ---- Krylov accelerators around the cycle ------------------------------------------------------------ */
/* loops over levels as vecloop.ct / matloop.ct do: ALL_VECTORS = every row of bl..level; ON_SURFACE = FINE_GRID_DOF rows of
 * fullrefinelevel..level-1 and NEW_DEFECT rows of level */
#define ALL_LOOP(stmt) for (int l = bl; l <= level; l++) { const ugport_level *L = &lv[l]; const int rm = 0; (void)rm; stmt; }
#define SURF_LOOP(stmt) for (int l = fr; l <= level; l++) { const ugport_level *L = &lv[l]; const int rm = l < level ? 2 : 1; stmt; }

static double surf_ddot(const ugport_level *lv, int fr, int level, double **x, double **y)
{
  double s = 0.0;
  SURF_LOOP(ugport_ddot_acc(L, rm, x[l], y[l], &s));
  return s;
}

/* ddotw ugblas.cc:3023-3045: per-component sums, then *s = 0; *s += w[i]*a[i] */
static double surf_ddotw(const ugport_level *lv, int fr, int level, double **x, double **y, const double *w)
{
  double a[UGPORT_MAX_BS] = {0.0, 0.0, 0.0}, s = 0.0;
  SURF_LOOP(ugport_ddotx_acc(L, rm, x[l], y[l], a));
  for (int i = 0; i < lv[level].bs; i++) s += w[i] * a[i];
  return s;
}

/* LinearSolver ls.cc:637-749 with Prepare/Update/Close = CGPrepare :976, CGUpdate :989-1027, CGClose :1159 (class `cg`).
 * p, tt: work vectors np->p, np->t; t: the cycle's temporary.  Returns the iterations done, -1 on lambda == 0. */
int ugport_cg_solve(const ugport_level *lv, const ugport_cfg *cfg, int fr, int level, double **x, double **b, double **c, double **t,
                    double **p, double **tt, int maxiter, const double *abslimit, const double *reduction, double *first_defect, double *history)
{
  int bs = lv[level].bs, bl = cfg->baselevel, it, done = 0;
  double last[UGPORT_MAX_BS], reach[UGPORT_MAX_BS];
  double *lu = cfg->base_hook ? NULL : ugport_base_factor(&lv[bl]);
  ALL_LOOP(ugport_dset(L, 0, p[l], 0.0));                                        /* CGPrepare */
  double rho = 1.0, lambda;
  ugport_ls_residuum(lv, fr, bl, level, b, last);
  for (int i = 0; i < bs; i++) { first_defect[i] = last[i]; reach[i] = last[i] * reduction[i]; if (reach[i] == 0.0) reach[i] = reduction[i]; }
  if (sc_cmp(last, abslimit, bs)) { ugport_base_free(lu); return 0; }
  for (it = 0; it < maxiter; it++) {
    ugport_dset(&lv[level], 0, c[level], 0.0);
    if (ugport_lmgc(lv, cfg, lu, level, c, b, t)) { ugport_base_free(lu); return -1; }
    ALL_LOOP(ugport_dmatmul(L, 0, 0, tt[l], c[l]));                              /* t = A c         :1003 */
    ALL_LOOP(ugport_dadd(L, 0, b[l], tt[l]));                                    /* b += t          :1005 */
    lambda = surf_ddot(lv, fr, level, c, b);                                     /* (c,b)           :1007 */
    ALL_LOOP(ugport_dscal(L, 0, p[l], lambda / rho));                            /* p *= lambda/rho :1009 */
    rho = lambda;
    ALL_LOOP(ugport_dadd(L, 0, p[l], c[l]));                                     /* p += c          :1012 */
    ALL_LOOP(ugport_dmatmul(L, 0, 0, tt[l], p[l]));                              /* t = A p         :1014 */
    lambda = surf_ddot(lv, fr, level, tt, p);                                    /* (t,p)           :1016 */
    if (lambda == 0.0) { ugport_base_free(lu); return -1; }
    ALL_LOOP(ugport_daxpy(L, 0, x[l], rho / lambda, p[l]));                      /* :1019 */
    ALL_LOOP(ugport_daxpy(L, 0, b[l], -rho / lambda, tt[l]));                    /* :1021 */
    ugport_ls_residuum(lv, fr, bl, level, b, last);
    if (history) for (int i = 0; i < bs; i++) history[it * bs + i] = last[i];
    done = it + 1;
    if (sc_cmp(last, abslimit, bs) || sc_cmp(last, reach, bs)) break;
  }
  ugport_base_free(lu);
  return done;
}

/* BCGSSolver ls.cc:1864-2062 (class `bcgs`) with Iter = the cycle, B = NULL, restart = 0.  w: the SQUARED weights
 * (BCGSInit :1757).  work = r p v s t q.  history receives last_defect after every pass of the loop (its second
 * residuum, or the first one if the loop ended there).  Returns number_of_linear_iterations (two per full pass). */
int ugport_bcgs_solve(const ugport_level *lv, const ugport_cfg *cfg, int fr, int level, double **x, double **b, double **t,
                      double **r, double **p, double **v, double **s, double **tt, double **q, const double *w,
                      int maxiter, const double *abslimit, const double *reduction, double *first_defect, double *history)
{
  int bs = lv[level].bs, bl = cfg->baselevel, nit = 0, eq_count = 0, restart = 1;
  double last[UGPORT_MAX_BS], reach[UGPORT_MAX_BS], old[UGPORT_MAX_BS] = {-1.0, -1.0, -1.0};
  double alpha = 0.0, rho_new = 0.0, beta = 0.0, tsq = 0.0, rho = 0.0, omega = 0.0;
  double *lu = cfg->base_hook ? NULL : ugport_base_factor(&lv[bl]);
  ugport_ls_residuum(lv, fr, bl, level, b, last);
  for (int i = 0; i < bs; i++) { first_defect[i] = last[i]; reach[i] = last[i] * reduction[i]; if (reach[i] == 0.0) reach[i] = reduction[i]; }
  int converged = sc_cmp(last, abslimit, bs);
  for (int i = 0; i < maxiter; i++) {
    if (converged) break;
    if (restart) {
      ALL_LOOP(ugport_dset(L, 0, p[l], 0.0)); ALL_LOOP(ugport_dset(L, 0, v[l], 0.0)); ALL_LOOP(ugport_dcopy(L, 0, r[l], b[l]));
      alpha = rho = omega = 1.0;
      restart = 0;
    }
    rho_new = surf_ddotw(lv, fr, level, b, r, w);                                 /* :1934 */
    if (rho != 0.0 && omega != 0.0) beta = rho_new * alpha / rho / omega;
    ALL_LOOP(ugport_dscal(L, 0, p[l], beta));
    ALL_LOOP(ugport_dadd(L, 0, p[l], b[l]));
    ALL_LOOP(ugport_daxpy(L, 0, p[l], -beta * omega, v[l]));
    ALL_LOOP(ugport_dset(L, 0, q[l], 0.0));
    ALL_LOOP(ugport_dcopy(L, 0, s[l], p[l]));
    if (ugport_lmgc(lv, cfg, lu, level, q, p, t)) { ugport_base_free(lu); return -1; }     /* Iter(q, p) :1944 */
    ALL_LOOP(ugport_dcopy(L, 0, p[l], s[l]));
    SURF_LOOP(ugport_dmatmul(L, 0, rm, v[l], q[l]));                              /* v = A q, ON_SURFACE :1946 */
    alpha = surf_ddotw(lv, fr, level, v, r, w);
    if (alpha != 0.0) alpha = rho_new / alpha;
    ALL_LOOP(ugport_daxpy(L, 0, x[l], alpha, q[l]));
    nit++;
    ALL_LOOP(ugport_dcopy(L, 0, s[l], b[l]));
    ALL_LOOP(ugport_daxpy(L, 0, s[l], -alpha, v[l]));
    ugport_ls_residuum(lv, fr, bl, level, s, last);                               /* :1975 */
    if (sc_cmp(last, abslimit, bs) || sc_cmp(last, reach, bs)) {
      ALL_LOOP(ugport_dcopy(L, 0, b[l], s[l]));
      converged = 1;
      if (history) for (int k = 0; k < bs; k++) history[i * bs + k] = last[k];
      break;
    }
    ALL_LOOP(ugport_dset(L, 0, q[l], 0.0));
    ALL_LOOP(ugport_dcopy(L, 0, tt[l], s[l]));
    if (ugport_lmgc(lv, cfg, lu, level, q, s, t)) { ugport_base_free(lu); return -1; }     /* Iter(q, s) :1991 */
    ALL_LOOP(ugport_dcopy(L, 0, s[l], tt[l]));
    SURF_LOOP(ugport_dmatmul(L, 0, rm, tt[l], q[l]));                             /* t = A q :2003 */
    tsq = surf_ddotw(lv, fr, level, tt, tt, w);
    omega = surf_ddotw(lv, fr, level, s, tt, w);
    if (tsq != 0.0) omega /= tsq;
    ALL_LOOP(ugport_daxpy(L, 0, x[l], omega, q[l]));
    ALL_LOOP(ugport_dcopy(L, 0, b[l], s[l]));
    ALL_LOOP(ugport_daxpy(L, 0, b[l], -omega, tt[l]));
    rho = rho_new;
    ugport_ls_residuum(lv, fr, bl, level, b, last);
    if (history) for (int k = 0; k < bs; k++) history[i * bs + k] = last[k];
    nit++;
    if (sc_cmp(last, abslimit, bs) || sc_cmp(last, reach, bs)) { converged = 1; break; }
    int eq = 1;                                                                   /* sc_eq npscan.cc:1095, ac = 1e-4 */
    for (int k = 0; k < bs; k++) if (last[k] < 0.0 || old[k] < 0.0 || fabs(last[k] - old[k]) > 1e-4 * sqrt(last[k] * old[k])) eq = 0;
    eq_count = eq ? eq_count + 1 : 0;
    for (int k = 0; k < bs; k++) old[k] = last[k];
    if (eq_count > 4) break;
  }
  ugport_base_free(lu);
  return nit;
}

/* ------------------------------------------------------------------------------------------------------------------------------------
 * Element-loop assembly: one level of LocalAssemble np/procs/assemble.cc:671-697 + that level's share of
 * NPLocalAssemblePostMatrix :624 (AssembleDirichletBoundary np/udm/disctools.cc:1837), sequential scatter loop like the reference,
 * with the element kernel of oracle/ug_driver.cc's class `fe` (UG leaves AssembleLocal to the application): simplices with the
 * centroid rule, tensor elements with 2-point Gauss, local matrix summed over the quadrature points, then added once per element. */
static double fe_det_inv(int dim, double J[3][3], double Ji[3][3])
{
  if (dim == 2) {
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    return det;
  }
  double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
}

typedef struct { double w, N[8], G[8][3]; } fe_qpt;

static int fe_element_qps(int dim, int nc, double X[8][3], fe_qpt *qp)
{
  static const int LOC[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};   /* UG's corner numbering */
  double J[3][3], Ji[3][3];
  if (nc == dim + 1) {
    for (int d = 0; d < dim; d++) for (int k = 0; k < dim; k++) J[k][d] = X[k + 1][d] - X[0][d];
    double det = fe_det_inv(dim, J, Ji);
    qp[0].w = fabs(det) / ((dim == 3) ? 6.0 : 2.0);
    for (int i = 0; i < nc; i++) qp[0].N[i] = 1.0 / nc;
    for (int d = 0; d < dim; d++) {
      double s = 0;
      for (int k = 0; k < dim; k++) { qp[0].G[k + 1][d] = Ji[d][k]; s += Ji[d][k]; }
      qp[0].G[0][d] = -s;
    }
    return 1;
  }
  const double g[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
  int nq = 0;
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < ((dim == 3) ? 2 : 1); c++) {
    double xi[3] = {g[a], g[b], g[c]}, dN[8][3];
    for (int i = 0; i < nc; i++) {
      double f[3], df[3];
      for (int d = 0; d < dim; d++) { f[d] = LOC[i][d] ? xi[d] : 1.0 - xi[d]; df[d] = LOC[i][d] ? 1.0 : -1.0; }
      qp[nq].N[i] = 1.0;
      for (int d = 0; d < dim; d++) qp[nq].N[i] *= f[d];
      for (int d = 0; d < dim; d++) { dN[i][d] = df[d]; for (int e = 0; e < dim; e++) if (e != d) dN[i][d] *= f[e]; }
    }
    for (int k = 0; k < dim; k++) for (int d = 0; d < dim; d++) { J[k][d] = 0; for (int i = 0; i < nc; i++) J[k][d] += dN[i][k] * X[i][d]; }
    double det = fe_det_inv(dim, J, Ji);
    qp[nq].w = fabs(det) / ((dim == 3) ? 8.0 : 4.0);
    for (int i = 0; i < nc; i++) for (int d = 0; d < dim; d++) { double s = 0; for (int k = 0; k < dim; k++) s += Ji[d][k] * dN[i][k]; qp[nq].G[i][d] = s; }
    nq++;
  }
  return nq;
}

int ugport_assemble(const ugport_level *L, const ugport_fe *fe, int64_t nelem, const int64_t *elem_ptr, const int32_t *elem_row,
                    const double *coef, const double *coord, const uint32_t *skip, const double *x, double *val, double *b)
{
  const int n = L->n, bs = L->bs, bb = bs * bs, dim = fe->dim;
  const double lam = fe->E * fe->nu / ((1 + fe->nu) * (1 - 2 * fe->nu)), mu = fe->E / (2 * (1 + fe->nu));
  for (int64_t i = 0; i < (int64_t)n * bs; i++) b[i] = 0.0;                           /* dset(b, 0) */
  for (int64_t i = 0; i < (int64_t)L->rowptr[n] * bb; i++) val[i] = 0.0;             /* dmatset(A, 0) */
  for (int64_t e = 0; e < nelem; e++) {
    const int32_t *er = elem_row + elem_ptr[e];
    const int nc = (int)(elem_ptr[e + 1] - elem_ptr[e]), m = nc * bs;
    if (nc != dim + 1 && nc != (1 << dim)) return 1;
    double X[8][3], def[24], mat[24 * 24];
    fe_qpt qp[8];
    for (int i = 0; i < nc; i++) for (int d = 0; d < dim; d++) X[i][d] = coord[(size_t)er[i] * dim + d];
    const int nq = fe_element_qps(dim, nc, X, qp);
    const double kappa = coef ? coef[e] : 1.0;
    for (int i = 0; i < m; i++) def[i] = 0.0;
    for (int i = 0; i < m * m; i++) mat[i] = 0.0;
    for (int q = 0; q < nq; q++) {
      const double wk = kappa * qp[q].w;
      for (int i = 0; i < nc; i++) {
        const double wn = qp[q].w * qp[q].N[i];
        for (int a = 0; a < bs; a++) def[i * bs + a] += wn * fe->source[a];
      }
      for (int i = 0; i < nc; i++)
        for (int j = 0; j < nc; j++) {
          const double *gi = qp[q].G[i], *gj = qp[q].G[j];
          double dot = 0; for (int d = 0; d < dim; d++) dot += gi[d] * gj[d];
          if (bs == 1) mat[i * m + j] += wk * dot;
          else
            for (int a = 0; a < bs; a++) for (int c = 0; c < bs; c++) {
              double k = lam * gi[a] * gj[c] + mu * gi[c] * gj[a] + ((a == c) ? mu * dot : 0.0);
              mat[(i * bs + a) * m + j * bs + c] += wk * k;
            }
        }
    }
    for (int i = 0; i < nc; i++) for (int a = 0; a < bs; a++) b[(size_t)er[i] * bs + a] += def[i * bs + a];
    for (int i = 0; i < nc; i++)
      for (int j = 0; j < nc; j++) {
        int64_t t = -1;
        for (int64_t u = L->rowptr[er[i]]; u < L->rowptr[er[i] + 1]; u++) if (L->col[u] == er[j]) { t = u; break; }
        if (t < 0) return 3;                                                         /* GetElementVVMPtrs: -3 */
        for (int a = 0; a < bs; a++) for (int c = 0; c < bs; c++) val[t * bb + a * bs + c] += mat[(i * bs + a) * m + j * bs + c];
      }
  }
  for (int r = 0; r < n; r++)                                                         /* AssembleDirichletBoundary */
    for (int a = 0; a < bs; a++)
      if (skip && (skip[r] & (1u << a))) {
        b[(size_t)r * bs + a] = x[(size_t)r * bs + a];
        for (int64_t u = L->rowptr[r]; u < L->rowptr[r + 1]; u++)
          for (int c = 0; c < bs; c++) val[u * bb + a * bs + c] = (u == L->rowptr[r] && c == a) ? 1.0 : 0.0;
      }
  return 0;
}
