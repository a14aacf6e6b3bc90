// tests/standin/uggpu_standin.cc -- TEST INFRASTRUCTURE ONLY.  Never shipped, never loaded by the product path.
//
// A CPU stand-in for libuggpu.so that exports the entry points the host numprocs bind (ug_b200/host/gpuls_np.cc UGGPU_FUNCS) and
// answers them with the oracle's plain-C restatement (oracle/ugport.c).  Its only purpose: the HOST logic of the gpuls numproc family
// -- flattening, level numbering (algebraic levels below 0), upload caching, the PreProcess / PostProcess brackets, the base-solver
// callback, LRESULT handling -- can be exercised inside the unmodified reference on a machine without a GPU
// (tests/test_host_numprocs.py, `-m "not gpu"`).  The GPU tests run the same drop-in cases against the real library.
//
// Build: g++ -O1 -shared -fPIC -I include -I oracle tests/standin/uggpu_standin.cc oracle/ugport.c -o tests/standin/libuggpu_standin.so
// (tests/test_host_numprocs.py does it).  bcgs restarts and partitions are not offered (calls fail).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "uggpu.h"
extern "C" {
#include "ugport.h"
}

namespace {
thread_local std::string g_err;
int fail(int code, const char *fmt, ...)
{
  char b[512];
  va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
  g_err = b;
  return code;
}
}  // namespace

// the host halves of the algebraic-level setup are the PRODUCT's own code (no device in them): the stand-in runs the very same coarsening
int uggpu_fail(int code, const char *fmt, ...)
{
  char b[512];
  va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
  g_err = b;
  return code;
}
#include "../../ug_b200/csrc/amg_host.inc"

namespace {

struct Lev {
  bool exists = false;
  int n = 0, bs = 0, mode = 0;
  std::vector<int32_t> rowptr, col, p_rowptr, p_col, r_rowptr, r_col;
  std::vector<double> p_w, r_w;
  std::vector<uint8_t> vclass, vnclass, ctl;
  std::vector<uint32_t> skip;
  std::map<int, std::vector<double> > mat;      // values per matrix handle (one pattern per level)
  std::map<int, std::vector<double> > vec;
};
}  // namespace

struct uggpu_ctx {
  Lev lev[UGGPU_MAX_LEVELS];
  int fullrefinelevel = 0;
};

namespace {
Lev *level_of(uggpu_ctx *c, int l)
{
  if (!c || l < 0 || l >= UGGPU_MAX_LEVELS || !c->lev[l].exists) { fail(UGGPU_ERROR, "level %d does not exist", l); return nullptr; }
  return &c->lev[l];
}
double *vec_of(uggpu_ctx *c, int l, int h, bool create)
{
  Lev *L = level_of(c, l);
  if (!L) return nullptr;
  auto it = L->vec.find(h);
  if (it == L->vec.end()) {
    if (!create) { fail(UGGPU_DESC_MISMATCH, "vector %d does not exist on level %d", h, l); return nullptr; }
    it = L->vec.emplace(h, std::vector<double>((size_t)L->n * L->bs + 1, 0.0)).first;
  }
  return it->second.data();
}
// the port's view of the levels with matrix handle A (ilu: values of handle Lh where they exist)
void port_levels(uggpu_ctx *c, int A, int Lh, std::vector<ugport_level> &out)
{
  out.assign(UGGPU_MAX_LEVELS, ugport_level());
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) {
    Lev &L = c->lev[l];
    if (!L.exists) continue;
    ugport_level &P = out[l];
    P.n = L.n; P.bs = L.bs;
    P.rowptr = L.rowptr.data(); P.col = L.col.data();
    auto it = L.mat.find(A);
    P.val = it == L.mat.end() ? nullptr : it->second.data();
    P.vclass = L.vclass.data(); P.vnclass = L.vnclass.data(); P.ctl = L.ctl.data(); P.skip = L.skip.data();
    P.p_rowptr = L.p_rowptr.data(); P.p_col = L.p_col.data(); P.p_w = L.p_w.data();
    P.r_rowptr = L.r_rowptr.data(); P.r_col = L.r_col.data(); P.r_w = L.r_w.data();
    auto il = L.mat.find(Lh);
    P.ilu = (Lh > 0 && il != L.mat.end()) ? il->second.data() : nullptr;
  }
}
void vec_array(uggpu_ctx *c, int h, bool create, int lo, int hi, std::vector<double *> &out)
{
  out.assign(UGGPU_MAX_LEVELS, nullptr);
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++)
    if (c->lev[l].exists) {
      auto it = c->lev[l].vec.find(h);
      if (it != c->lev[l].vec.end()) out[l] = it->second.data();
      else if (create && l >= lo && l <= hi) out[l] = vec_of(c, l, h, true);
    }
}

struct Hook { const uggpu_lmgc_cfg *cfg; uggpu_ctx *ctx; int c, b, A; };
// the cycle may run on other vectors than the solver's (bcgs: Iter(q, p) and Iter(q, s)): the handles are found from the storage
int base_hook(void *user, int level, double *c, double *b)
{
  Hook *h = (Hook *)user;
  int ch = h->c, bh = h->b;
  for (auto &kv : h->ctx->lev[level].vec) {
    if (kv.second.data() == c) ch = kv.first;
    if (kv.second.data() == b) bh = kv.first;
  }
  return h->cfg->base_solver(h->cfg->base_user, h->ctx, level, ch, bh, h->A);
}
// transfer modes: the port takes one flag for all levels and "by-matrix at and below level k"
void port_cfg(uggpu_ctx *c, const uggpu_lmgc_cfg *g, int level, Hook *hk, ugport_cfg *p)
{
  memset(p, 0, sizeof *p);
  p->nu1 = g->nu1; p->nu2 = g->nu2; p->gamma = g->gamma; p->baselevel = g->baselevel;
  for (int i = 0; i < UGPORT_MAX_BS; i++) { p->smooth_damp[i] = g->smooth_damp[i]; p->cycle_damp[i] = g->cycle_damp[i]; }
  p->base_maxit = g->base_maxit; p->base_reduction = g->base_reduction; p->base_abslimit = g->base_abslimit;
  p->smoother = g->smoother;
  p->level_opt = g->level_opt;
  // by-matrix levels: all of them ($M), or the lowest ones (the algebraic levels of an AMG transfer); anything else cannot be expressed
  int below = g->baselevel;
  for (int l = g->baselevel + 1; l <= level; l++)
    if (c->lev[l].mode == UGGPU_TRANSFER_IMAT && below == l - 1) below = l;
  p->imat = (below == level && level > g->baselevel) ? 1 : 0;
  p->imat_below = p->imat ? 0 : (below > g->baselevel ? below : -1);
  if (g->base_solver) { p->base_hook = base_hook; p->base_user = hk; }
}
int ilu_ready(uggpu_ctx *c, const uggpu_lmgc_cfg *g, int level, int A)
{
  if (g->smoother != UGGPU_SM_ILU) return 0;
  for (int l = g->baselevel + 1; l <= level; l++) {
    if (c->lev[l].mat.count(g->smoother_L)) continue;
    if (uggpu_dmatcopy(c, l, l, UGGPU_ALL_VECTORS, g->smoother_L, A)) return UGGPU_ERROR;
    if (int rc = uggpu_l_ilubthdecomp(c, l, g->smoother_L, g->ilu_beta)) return rc;
  }
  return 0;
}
}  // namespace

extern "C" {

int uggpu_ctx_create(int, uggpu_ctx **out) { *out = new uggpu_ctx(); return 0; }
int uggpu_ctx_destroy(uggpu_ctx *ctx) { delete ctx; return 0; }
const char *uggpu_last_error(void) { return g_err.c_str(); }
int64_t uggpu_launch_count(uggpu_ctx *) { return 0; }
int uggpu_set_fullrefinelevel(uggpu_ctx *ctx, int level) { ctx->fullrefinelevel = level; return 0; }

int uggpu_level_create(uggpu_ctx *ctx, int level, int n, int bs)
{
  if (level < 0 || level >= UGGPU_MAX_LEVELS) return fail(UGGPU_ERROR, "level %d out of range", level);
  if (bs < 1 || bs > UGGPU_MAX_BS) return fail(UGGPU_ERROR, "block size %d not supported", bs);
  ctx->lev[level] = Lev();
  Lev &L = ctx->lev[level];
  L.exists = true; L.n = n; L.bs = bs;
  L.vclass.assign(n, 3); L.vnclass.assign(n, 0); L.ctl.assign(n, 1); L.skip.assign(n, 0);
  return 0;
}

int uggpu_level_set_flags(uggpu_ctx *ctx, int level, const uint8_t *vclass, const uint8_t *vnclass, const uint8_t *ctl, const uint32_t *skip)
{
  Lev *L = level_of(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (vclass) L->vclass.assign(vclass, vclass + L->n);
  if (vnclass) L->vnclass.assign(vnclass, vnclass + L->n);
  if (ctl) L->ctl.assign(ctl, ctl + L->n);
  if (skip) L->skip.assign(skip, skip + L->n);
  return 0;
}

int uggpu_level_get_flags(uggpu_ctx *ctx, int level, uint8_t *vclass, uint8_t *vnclass, uint8_t *ctl, uint32_t *skip)
{
  Lev *L = level_of(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (vclass) memcpy(vclass, L->vclass.data(), L->n);
  if (vnclass) memcpy(vnclass, L->vnclass.data(), L->n);
  if (ctl) memcpy(ctl, L->ctl.data(), L->n);
  if (skip) memcpy(skip, L->skip.data(), sizeof(uint32_t) * L->n);
  return 0;
}

int uggpu_mat_set_pattern(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col)
{
  Lev *L = level_of(ctx, level);
  if (!L) return UGGPU_ERROR;
  L->rowptr.assign(rowptr, rowptr + L->n + 1);
  L->col.assign(col, col + rowptr[L->n]);
  L->mat[mat].assign((size_t)rowptr[L->n] * L->bs * L->bs, 0.0);
  return 0;
}

int uggpu_mat_set(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col, const double *val)
{
  if (uggpu_mat_set_pattern(ctx, level, mat, rowptr, col)) return UGGPU_ERROR;
  Lev &L = ctx->lev[level];
  L.mat[mat].assign(val, val + (size_t)rowptr[L.n] * L.bs * L.bs);
  return 0;
}

int64_t uggpu_mat_nnz(uggpu_ctx *ctx, int level, int mat)
{
  Lev *L = level_of(ctx, level);
  return (L && L->mat.count(mat)) ? (int64_t)L->col.size() : -1;
}

int uggpu_mat_get(uggpu_ctx *ctx, int level, int mat, int32_t *rowptr, int32_t *col, double *val)
{
  Lev *L = level_of(ctx, level);
  if (!L || !L->mat.count(mat)) return fail(UGGPU_DESC_MISMATCH, "matrix %d does not exist on level %d", mat, level);
  if (rowptr) memcpy(rowptr, L->rowptr.data(), sizeof(int32_t) * L->rowptr.size());
  if (col) memcpy(col, L->col.data(), sizeof(int32_t) * L->col.size());
  if (val) memcpy(val, L->mat[mat].data(), sizeof(double) * L->mat[mat].size());
  return 0;
}

int uggpu_transfer_set(uggpu_ctx *ctx, int level, const int32_t *p_rowptr, const int32_t *p_col, const double *p_w, const int32_t *r_rowptr,
                       const int32_t *r_col, const double *r_w)
{
  Lev *L = level_of(ctx, level), *Lc = level_of(ctx, level - 1);
  if (!L || !Lc) return UGGPU_ERROR;
  L->p_rowptr.assign(p_rowptr, p_rowptr + L->n + 1);
  L->p_col.assign(p_col, p_col + p_rowptr[L->n]); L->p_w.assign(p_w, p_w + p_rowptr[L->n]);
  L->r_rowptr.assign(r_rowptr, r_rowptr + Lc->n + 1);
  L->r_col.assign(r_col, r_col + r_rowptr[Lc->n]); L->r_w.assign(r_w, r_w + r_rowptr[Lc->n]);
  L->mode = UGGPU_TRANSFER_STANDARD;
  return 0;
}

int uggpu_transfer_set_mode(uggpu_ctx *ctx, int level, int mode)
{
  Lev *L = level_of(ctx, level);
  if (!L) return UGGPU_ERROR;
  L->mode = mode;
  return 0;
}

int uggpu_vec_alloc(uggpu_ctx *ctx, int level, int vec) { return vec_of(ctx, level, vec, true) ? 0 : UGGPU_ERROR; }
int uggpu_vec_upload(uggpu_ctx *ctx, int level, int vec, const double *host)
{
  double *v = vec_of(ctx, level, vec, true);
  if (!v) return UGGPU_ERROR;
  memcpy(v, host, sizeof(double) * (size_t)ctx->lev[level].n * ctx->lev[level].bs);
  return 0;
}
int uggpu_vec_download(uggpu_ctx *ctx, int level, int vec, double *host)
{
  double *v = vec_of(ctx, level, vec, false);
  if (!v) return UGGPU_DESC_MISMATCH;
  memcpy(host, v, sizeof(double) * (size_t)ctx->lev[level].n * ctx->lev[level].bs);
  return 0;
}

int uggpu_gs_preprocess(uggpu_ctx *ctx, int level, int) { return level_of(ctx, level) ? 0 : UGGPU_ERROR; }

int uggpu_dmatcopy(uggpu_ctx *ctx, int fl, int tl, int mode, int M, int A)
{
  if (mode != UGGPU_ALL_VECTORS) return fail(UGGPU_ERROR, "dmatcopy: ALL_VECTORS only");
  for (int l = fl; l <= tl; l++) {
    Lev *L = level_of(ctx, l);
    if (!L || !L->mat.count(A)) return fail(UGGPU_DESC_MISMATCH, "matrix %d does not exist on level %d", A, l);
    L->mat[M] = L->mat[A];
  }
  return 0;
}

int uggpu_l_ilubthdecomp(uggpu_ctx *ctx, int level, int M, const double *beta)
{
  Lev *L = level_of(ctx, level);
  if (!L || !L->mat.count(M)) return fail(UGGPU_DESC_MISMATCH, "matrix %d does not exist on level %d", M, level);
  std::vector<ugport_level> lv;
  port_levels(ctx, M, 0, lv);
  std::vector<double> out(L->mat[M].size());
  if (int rc = ugport_ilu_decomp(&lv[level], beta, out.data())) return fail(UGGPU_SMALL_DIAG, "ilu decomposition failed (%d)", rc);
  L->mat[M] = out;
  return 0;
}

int uggpu_smooth(uggpu_ctx *ctx, int level, int kind, int x, int b, int A, const double *damp, int tmp)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, A, kind == UGGPU_SM_ILU ? tmp : 0, lv);
  double *xv = vec_of(ctx, level, x, true), *bv = vec_of(ctx, level, b, false);
  if (!xv || !bv || !lv[level].val) return UGGPU_DESC_MISMATCH;
  std::vector<double> t((size_t)lv[level].n * lv[level].bs + 1);
  if (int rc = ugport_smooth(&lv[level], kind, xv, bv, damp, t.data())) return fail(rc, "smoothing step failed (%d)", rc);
  return 0;
}
int uggpu_jac_smooth(uggpu_ctx *ctx, int level, int x, int b, int A, const double *damp) { return uggpu_smooth(ctx, level, UGGPU_SM_JAC, x, b, A, damp, 0); }

int uggpu_restrict(uggpu_ctx *ctx, int level, int to, int from, const double *damp)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, 0, 0, lv);
  if (!level_of(ctx, level) || !level_of(ctx, level - 1)) return UGGPU_ERROR;
  double *tv = vec_of(ctx, level - 1, to, true), *fv = vec_of(ctx, level, from, false);
  if (!tv || !fv) return UGGPU_DESC_MISMATCH;
  (ctx->lev[level].mode == UGGPU_TRANSFER_IMAT ? ugport_restrict_imat : ugport_restrict)(&lv[level], &lv[level - 1], tv, fv, damp);
  return 0;
}
int uggpu_interpolate_correction(uggpu_ctx *ctx, int level, int to, int from, const double *damp)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, 0, 0, lv);
  if (!level_of(ctx, level) || !level_of(ctx, level - 1)) return UGGPU_ERROR;
  double *tv = vec_of(ctx, level, to, true), *fv = vec_of(ctx, level - 1, from, false);
  if (!tv || !fv) return UGGPU_DESC_MISMATCH;
  (ctx->lev[level].mode == UGGPU_TRANSFER_IMAT ? ugport_interpolate_imat : ugport_interpolate)(&lv[level], &lv[level - 1], tv, fv, damp);
  return 0;
}

int uggpu_minimize_level(uggpu_ctx *ctx, int level, int c, int b, int A, int t)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, A, 0, lv);
  double *cv = vec_of(ctx, level, c, false), *bv = vec_of(ctx, level, b, false), *tv = vec_of(ctx, level, t, true);
  if (!cv || !bv || !tv || !lv[level].val) return UGGPU_DESC_MISMATCH;
  ugport_minimize_level(&lv[level], cv, bv, tv);
  return 0;
}

// level-1 from the interpolation rows: flags, by-matrix stencils, Galerkin matrix (pattern: ugport_galerkin_pattern, values: ugport_galerkin)
static int standin_build_level(uggpu_ctx *ctx, int level, int A, int nc, const std::vector<uint8_t> &cnclass, const std::vector<uint32_t> &cskip,
                               const std::vector<int32_t> &prp, const std::vector<int32_t> &pcol, const std::vector<double> &pw)
{
  Lev &F = ctx->lev[level];
  const int n = F.n;
  std::vector<uint8_t> cclass((size_t)nc, 3), cctl((size_t)nc, 1);
  if (uggpu_level_create(ctx, level - 1, nc, 1) || uggpu_level_set_flags(ctx, level - 1, cclass.data(), cnclass.data(), cctl.data(), cskip.data())) return UGGPU_ERROR;
  std::vector<int32_t> rrp((size_t)nc + 1, 0), rcol((size_t)prp[n] + 1);
  std::vector<double> rw((size_t)prp[n] + 1);
  for (int v = 0; v < n; v++) if (F.vclass[v] >= 2) for (int e = prp[v]; e < prp[v + 1]; e++) rrp[pcol[e] + 1]++;
  for (int k = 0; k < nc; k++) rrp[k + 1] += rrp[k];
  { std::vector<int32_t> fill(rrp.begin(), rrp.end() - 1);
    for (int v = 0; v < n; v++) if (F.vclass[v] >= 2) for (int e = prp[v]; e < prp[v + 1]; e++) { const int32_t pos = fill[pcol[e]]++; rcol[pos] = v; rw[pos] = pw[e]; } }
  if (uggpu_transfer_set(ctx, level, prp.data(), pcol.data(), pw.data(), rrp.data(), rcol.data(), rw.data())) return UGGPU_ERROR;
  uggpu_transfer_set_mode(ctx, level, UGGPU_TRANSFER_IMAT);
  std::vector<int32_t> crp((size_t)nc + 1), ccol;
  if (ugport_galerkin_pattern(n, nc, F.rowptr.data(), F.col.data(), prp.data(), pcol.data(), nullptr, nullptr, crp.data(), nullptr)) return fail(UGGPU_ERROR, "galerkin pattern");
  ccol.resize((size_t)crp[nc] + 1);
  if (ugport_galerkin_pattern(n, nc, F.rowptr.data(), F.col.data(), prp.data(), pcol.data(), nullptr, nullptr, crp.data(), ccol.data())) return fail(UGGPU_ERROR, "galerkin pattern");
  if (uggpu_mat_set_pattern(ctx, level - 1, A, crp.data(), ccol.data())) return UGGPU_ERROR;
  std::vector<ugport_level> lv;
  port_levels(ctx, A, 0, lv);
  if (ugport_galerkin(&lv[level], &lv[level - 1], F.mat[A].data(), ctx->lev[level - 1].mat[A].data())) return fail(UGGPU_ERROR, "galerkin product");
  return 0;
}

int uggpu_amg_coarsen_rs(uggpu_ctx *ctx, int level, int A, double theta, int *n_coarse)
{
  Lev *L = level_of(ctx, level);
  if (!L || level < 1 || L->bs != 1 || !L->mat.count(A)) return fail(UGGPU_ERROR, "uggpu_amg_coarsen_rs: bad level");
  const int n = L->n; const size_t nnz = L->col.size();
  std::vector<int32_t> prp((size_t)n + 1), pcol(nnz + (size_t)n + 1);
  std::vector<double> pw(nnz + (size_t)n + 1);
  std::vector<uint8_t> coarse((size_t)n + 1);
  int nc = 0;
  if (int rc = uggpu_amg_rs_host(n, L->rowptr.data(), L->col.data(), L->mat[A].data(), L->skip.data(), theta, coarse.data(), prp.data(), pcol.data(), pw.data(), &nc)) return rc;
  *n_coarse = nc;
  if (nc == 0 || nc == n) { *n_coarse = 0; return 0; }
  std::vector<uint8_t> cnclass((size_t)nc); std::vector<uint32_t> cskip((size_t)nc);
  for (int v = 0, k = 0; v < n; v++) if (coarse[v]) { cnclass[k] = L->vclass[v]; cskip[k] = L->skip[v]; k++; }
  return standin_build_level(ctx, level, A, nc, cnclass, cskip, prp, pcol, pw);
}

int uggpu_amg_coarsen_vanek(uggpu_ctx *ctx, int level, int A, double theta, int smooth, int *n_coarse)
{
  Lev *L = level_of(ctx, level);
  if (!L || level < 1 || L->bs != 1 || !L->mat.count(A)) return fail(UGGPU_ERROR, "uggpu_amg_coarsen_vanek: bad level");
  const int n = L->n; const size_t nnz = L->col.size();
  std::vector<int32_t> prp((size_t)n + 1), pcol(nnz + (size_t)n + 1), cluster((size_t)n + 1), seed((size_t)n + 1);
  std::vector<double> pw(nnz + (size_t)n + 1);
  int nc = 0;
  if (int rc = uggpu_amg_vanek_host(n, L->rowptr.data(), L->col.data(), L->mat[A].data(), L->skip.data(), theta, smooth, cluster.data(), seed.data(), prp.data(), pcol.data(), pw.data(), &nc)) return rc;
  *n_coarse = nc;
  if (nc == 0 || nc == n) { *n_coarse = 0; return 0; }
  std::vector<uint8_t> cnclass((size_t)nc); std::vector<uint32_t> cskip((size_t)nc, 0u);
  for (int c = 0; c < nc; c++) cnclass[c] = L->vclass[seed[c]];
  return standin_build_level(ctx, level, A, nc, cnclass, cskip, prp, pcol, pw);
}

int uggpu_lmgc_preprocess(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int A)
{
  for (int l = cfg->baselevel; l <= level; l++) if (!level_of(ctx, l)) return UGGPU_ERROR;
  return ilu_ready(ctx, cfg, level, A);
}

int uggpu_lmgc(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A)
{
  if (int rc = ilu_ready(ctx, cfg, level, A)) return rc;
  std::vector<ugport_level> lv;
  port_levels(ctx, A, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : 0, lv);
  Hook hk = {cfg, ctx, c, b, A};
  ugport_cfg p;
  port_cfg(ctx, cfg, level, &hk, &p);
  std::vector<double *> cv, bv, tv;
  vec_array(ctx, c, true, cfg->baselevel, level, cv); vec_array(ctx, b, true, cfg->baselevel, level, bv); vec_array(ctx, cfg->t, true, cfg->baselevel, level, tv);
  double *lu = p.base_hook ? nullptr : ugport_base_factor(&lv[cfg->baselevel]);
  const int rc = ugport_lmgc(lv.data(), &p, lu, level, cv.data(), bv.data(), tv.data());
  ugport_base_free(lu);
  return rc ? fail(UGGPU_ERROR, "cycle failed (%d)", rc) : 0;
}

int uggpu_ls_defect(uggpu_ctx *ctx, int bl, int level, int x, int b, int A)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, A, 0, lv);
  std::vector<double *> xv, bv;
  vec_array(ctx, x, false, 0, 0, xv); vec_array(ctx, b, false, 0, 0, bv);
  const int fr = ctx->fullrefinelevel < level ? ctx->fullrefinelevel : level;
  for (int l = fr; l <= level; l++) if (!xv[l] || !bv[l] || !lv[l].val) return fail(UGGPU_DESC_MISMATCH, "ls_defect: level %d incomplete", l);
  ugport_ls_defect(lv.data(), fr, bl, level, xv.data(), bv.data());
  return 0;
}

int uggpu_ls_residuum(uggpu_ctx *ctx, int bl, int level, int b, uggpu_lresult *res)
{
  std::vector<ugport_level> lv;
  port_levels(ctx, 0, 0, lv);
  std::vector<double *> bv;
  vec_array(ctx, b, false, 0, 0, bv);
  const int fr = ctx->fullrefinelevel < level ? ctx->fullrefinelevel : level;
  for (int l = fr; l <= level; l++) if (!bv[l]) return fail(UGGPU_DESC_MISMATCH, "ls_residuum: level %d incomplete", l);
  ugport_ls_residuum(lv.data(), fr, bl, level, bv.data(), res->last_defect);
  return 0;
}

static void finish(uggpu_lresult *res, int its, int bs, const double *first, const double *hist, const double *absl, const double *red)
{
  res->number_of_linear_iterations = its;
  for (int i = 0; i < bs; i++) { res->first_defect[i] = first[i]; if (its > 0) res->last_defect[i] = hist[(size_t)(its - 1) * bs + i]; else res->last_defect[i] = first[i]; }
  bool ca = true, cr = true;
  for (int i = 0; i < bs; i++) {
    if (!(fabs(res->last_defect[i]) < fabs(absl[i]))) ca = false;
    double reach = first[i] * red[i]; if (reach == 0.0) reach = red[i];
    if (!(fabs(res->last_defect[i]) < fabs(reach))) cr = false;
  }
  res->converged = (ca || cr) ? 1 : 0;
  res->error_code = 0;
}

int uggpu_ls_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, int c, int maxiter, const double *abslimit,
                   const double *reduction, uggpu_lresult *res, double *history)
{
  (void)bl;
  if (int rc = ilu_ready(ctx, cfg, level, A)) return rc;
  std::vector<ugport_level> lv;
  port_levels(ctx, A, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : 0, lv);
  Hook hk = {cfg, ctx, c, b, A};
  ugport_cfg p;
  port_cfg(ctx, cfg, level, &hk, &p);
  std::vector<double *> xv, bv, cv, tv;
  vec_array(ctx, x, false, 0, 0, xv); vec_array(ctx, b, false, 0, 0, bv);
  vec_array(ctx, c, true, cfg->baselevel, level, cv); vec_array(ctx, cfg->t, true, cfg->baselevel, level, tv);
  const int bs = ctx->lev[level].bs;
  std::vector<double> hist((size_t)(maxiter > 0 ? maxiter : 1) * bs, 0.0);
  double first[UGPORT_MAX_BS];
  const int fr = ctx->fullrefinelevel < level ? ctx->fullrefinelevel : level;
  const int its = ugport_solve(lv.data(), &p, fr, level, xv.data(), bv.data(), cv.data(), tv.data(), maxiter, abslimit, reduction, first, hist.data());
  if (its < 0) return fail(UGGPU_ERROR, "solve failed");
  if (history) memcpy(history, hist.data(), sizeof(double) * (size_t)its * bs);
  finish(res, its, bs, first, hist.data(), abslimit, reduction);
  return 0;
}

int uggpu_cg_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, int c, int pp, int t, int maxiter,
                   const double *abslimit, const double *reduction, uggpu_lresult *res, double *history)
{
  (void)bl;
  if (int rc = ilu_ready(ctx, cfg, level, A)) return rc;
  std::vector<ugport_level> lv;
  port_levels(ctx, A, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : 0, lv);
  Hook hk = {cfg, ctx, c, b, A};
  ugport_cfg p;
  port_cfg(ctx, cfg, level, &hk, &p);
  std::vector<double *> xv, bv, cv, tv, pv, ttv;
  vec_array(ctx, x, false, 0, 0, xv); vec_array(ctx, b, false, 0, 0, bv);
  vec_array(ctx, c, true, cfg->baselevel, level, cv); vec_array(ctx, cfg->t, true, cfg->baselevel, level, tv);
  vec_array(ctx, pp, true, cfg->baselevel, level, pv); vec_array(ctx, t, true, cfg->baselevel, level, ttv);
  const int bs = ctx->lev[level].bs;
  std::vector<double> hist((size_t)(maxiter > 0 ? maxiter : 1) * bs, 0.0);
  double first[UGPORT_MAX_BS];
  const int fr = ctx->fullrefinelevel < level ? ctx->fullrefinelevel : level;
  const int its = ugport_cg_solve(lv.data(), &p, fr, level, xv.data(), bv.data(), cv.data(), tv.data(), pv.data(), ttv.data(), maxiter, abslimit, reduction, first, hist.data());
  if (its < 0) return fail(UGGPU_ERROR, "cg failed");
  if (history) memcpy(history, hist.data(), sizeof(double) * (size_t)its * bs);
  finish(res, its, bs, first, hist.data(), abslimit, reduction);
  return 0;
}

int uggpu_bcgs_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, const int *work, const double *weight, int restart_every,
                     int maxiter, const double *abslimit, const double *reduction, uggpu_lresult *res, double *history)
{
  (void)bl;
  if (restart_every != 0) return fail(UGGPU_ERROR, "the CPU stand-in offers bcgs without restarts");
  if (int rc = ilu_ready(ctx, cfg, level, A)) return rc;
  std::vector<ugport_level> lv;
  port_levels(ctx, A, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : 0, lv);
  Hook hk = {cfg, ctx, work[5], work[1], A};
  ugport_cfg p;
  port_cfg(ctx, cfg, level, &hk, &p);
  std::vector<double *> xv, bv, tv, wv[6];
  vec_array(ctx, x, false, 0, 0, xv); vec_array(ctx, b, false, 0, 0, bv);
  vec_array(ctx, cfg->t, true, cfg->baselevel, level, tv);
  for (int i = 0; i < 6; i++) vec_array(ctx, work[i], true, cfg->baselevel, level, wv[i]);      // r p v s t q
  const int bs = ctx->lev[level].bs;
  double w2[UGPORT_MAX_BS];
  for (int i = 0; i < UGPORT_MAX_BS; i++) w2[i] = weight[i] * weight[i];                         // BCGSInit ls.cc:1757
  std::vector<double> hist((size_t)(maxiter > 0 ? maxiter : 1) * bs, 0.0);
  double first[UGPORT_MAX_BS];
  const int fr = ctx->fullrefinelevel < level ? ctx->fullrefinelevel : level;
  const int its = ugport_bcgs_solve(lv.data(), &p, fr, level, xv.data(), bv.data(), tv.data(), wv[0].data(), wv[1].data(), wv[2].data(), wv[3].data(), wv[4].data(),
                                    wv[5].data(), w2, maxiter, abslimit, reduction, first, hist.data());
  if (its < 0) return fail(UGGPU_ERROR, "bcgs failed");
  const int nhist = (its + 1) / 2;
  if (history) memcpy(history, hist.data(), sizeof(double) * (size_t)nhist * bs);
  finish(res, nhist, bs, first, hist.data(), abslimit, reduction);
  res->number_of_linear_iterations = its;
  return 0;
}
int uggpu_assemble(uggpu_ctx *ctx, int level, int x, int b, int A, const uggpu_fe_cfg *cfg, int64_t nelem, const int64_t *elem_ptr, const int32_t *elem_row,
                   const double *coef, const double *coord, const uint32_t *skip)
{
  Lev *L = level_of(ctx, level);
  if (!L || !cfg || !L->mat.count(A)) return fail(UGGPU_DESC_MISMATCH, "uggpu_assemble: level or matrix missing");
  double *xv = vec_of(ctx, level, x, false), *bv = vec_of(ctx, level, b, true);
  if (!xv || !bv) return UGGPU_DESC_MISMATCH;
  std::vector<uint32_t> sk((size_t)L->n + 1, 0u);
  if (skip) memcpy(sk.data(), skip, sizeof(uint32_t) * (size_t)L->n);
  std::vector<double> cf;
  if (!coef) { cf.assign((size_t)nelem + 1, 1.0); coef = cf.data(); }
  std::vector<ugport_level> lv;
  port_levels(ctx, A, 0, lv);
  ugport_fe fe;
  fe.problem = cfg->problem; fe.dim = cfg->dim; fe.E = cfg->E; fe.nu = cfg->nu;
  for (int i = 0; i < UGPORT_MAX_BS; i++) fe.source[i] = cfg->source[i];
  const int rc = ugport_assemble(&lv[level], &fe, nelem, elem_ptr, elem_row, coef, coord, sk.data(), xv, L->mat[A].data(), bv);
  if (rc) return fail(rc == 3 ? UGGPU_DESC_MISMATCH : UGGPU_ERROR, "uggpu_assemble: the restatement returned %d", rc);
  L->skip.assign(sk.begin(), sk.begin() + L->n);                    // VECSKIP := skip (SetElementDirichletFlags)
  return 0;
}

// savedata / loaddata: the FORMAT halves are host-only functions of the product library (uggpu_data_write / uggpu_data_read, no device in
// them); the stand-in borrows them from libuggpu.so next to the repository's ug_b200/lib and does the gather / scatter itself
}  // extern "C"
#include <dlfcn.h>
#include <cstdlib>
#include <climits>
namespace {
typedef int (*data_write_fn)(const char *, const char *, const uggpu_data_general *, int, const int *, const char *const *, const char *const *, int64_t, const double *);
typedef int (*data_read_fn)(const char *, uggpu_data_general *, int *, int *, int, int64_t *, double *, int64_t);
data_write_fn p_write = nullptr;
data_read_fn p_read = nullptr;
int bind_format_functions()
{
  if (p_write && p_read) return 0;
  Dl_info info;
  if (!dladdr((void *)&bind_format_functions, &info) || !info.dli_fname) return 1;
  char real[4096];
  if (!realpath(info.dli_fname, real)) return 1;
  std::string dir = real;                                            // .../tests/standin/libuggpu_standin.so
  for (int k = 0; k < 3; k++) { size_t pos = dir.find_last_of('/'); if (pos == std::string::npos) return 1; dir.erase(pos); }
  void *h = dlopen((dir + "/ug_b200/lib/libuggpu.so").c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  p_write = (data_write_fn)dlsym(h, "uggpu_data_write");
  p_read = (data_read_fn)dlsym(h, "uggpu_data_read");
  return (p_write && p_read) ? 0 : 1;
}
}  // namespace
extern "C" {

int uggpu_savedata(uggpu_ctx *ctx, const char *filename, const char *type, const uggpu_data_general *g, int nvd, const int *vec, const char *const *vdname,
                   const char *const *compnames, int64_t nnode, const int32_t *id_level, const int32_t *id_row)
{
  if (bind_format_functions()) return fail(UGGPU_ERROR, "the CPU stand-in cannot bind uggpu_data_write / uggpu_data_read of libuggpu.so");
  std::vector<int> ncomp(nvd);
  int sum = 0;
  for (int i = 0; i < nvd; i++) { ncomp[i] = (int)strlen(compnames[i]); sum += ncomp[i]; }
  std::vector<double> body((size_t)nnode * sum + 1);
  for (int64_t id = 0; id < nnode; id++) {
    int o = 0;
    for (int i = 0; i < nvd; i++) {
      const double *v = vec_of(ctx, id_level[id], vec[i], false);
      if (!v || ctx->lev[id_level[id]].bs != ncomp[i]) return fail(UGGPU_DESC_MISMATCH, "uggpu_savedata: vector %d on level %d", vec[i], id_level[id]);
      for (int c = 0; c < ncomp[i]; c++) body[(size_t)id * sum + o + c] = v[(size_t)id_row[id] * ncomp[i] + c];
      o += ncomp[i];
    }
  }
  return p_write(filename, type, g, nvd, ncomp.data(), vdname, compnames, nnode, body.data());
}

int uggpu_loaddata(uggpu_ctx *ctx, const char *filename, int nvd, const int *vec, int64_t nnode, const int32_t *id_level, const int32_t *id_row, uggpu_data_general *general_out)
{
  if (bind_format_functions()) return fail(UGGPU_ERROR, "the CPU stand-in cannot bind uggpu_data_write / uggpu_data_read of libuggpu.so");
  uggpu_data_general g;
  int fnvd = 0, ncomp[64];
  int64_t ndata = 0;
  if (p_read(filename, &g, &fnvd, ncomp, 64, &ndata, nullptr, 0)) return fail(UGGPU_ERROR, "uggpu_loaddata: cannot read the header of %s", filename);
  int sum = 0;
  for (int i = 0; i < fnvd; i++) sum += ncomp[i];
  std::vector<double> body((size_t)ndata + 1);
  if (p_read(filename, &g, &fnvd, ncomp, 64, &ndata, body.data(), ndata)) return fail(UGGPU_ERROR, "uggpu_loaddata: cannot read %s", filename);
  if (sum <= 0 || ndata != nnode * sum) return fail(UGGPU_ERROR, "uggpu_loaddata: %s holds %lld values, expected %lld", filename, (long long)ndata, (long long)nnode * sum);
  for (int64_t id = 0; id < nnode; id++) {
    int o = 0;
    for (int i = 0; i < fnvd; i++) {
      if (i < nvd && vec[i] >= 0) {
        double *v = vec_of(ctx, id_level[id], vec[i], true);
        if (!v || ctx->lev[id_level[id]].bs != ncomp[i]) return fail(UGGPU_DESC_MISMATCH, "uggpu_loaddata: vector %d on level %d", vec[i], id_level[id]);
        for (int c = 0; c < ncomp[i]; c++) v[(size_t)id_row[id] * ncomp[i] + c] = body[(size_t)id * sum + o + c];
      }
      o += ncomp[i];
    }
  }
  if (general_out) *general_out = g;
  return 0;
}

}  // extern "C"
