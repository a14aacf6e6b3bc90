// gpuls_np.cc -- see gpuls_np.h.  Host-side product code of the `gpuls` numproc family (no CUDA, no oracle).
//
// Structure mirrors the reference classes one to one (file:line in the comments); the BLAS/iter/transfer calls of
// the CPU classes are replaced by calls through the C-ABI of include/uggpu.h on a device mirror of the multigrid:
//   * PreProcess flattens UG's VECTOR/MATRIX lists of the levels involved (gpuls_flatten.cc) and uploads them once;
//     PostProcess releases the device resources (the pattern of np/amglib/amg_ug.cc:207-390,599-615);
//   * every entry point is synchronous: its vector arguments are gathered from the VVALUEs and uploaded on entry,
//     results are downloaded and scattered back before it returns, because UG callers read VVALUEs immediately;
//   * linear_solver.gpuls + iter.gpulmgc keep the whole solve resident: x, b go up once, all iterations run on the
//     device (uggpu_ls_solve), x, b, c come back once.
#include "config.h"
#include "gpuls_np.h"
#include "gpuls_flatten.h"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "np.h"
#include "disctools.h"
#include "numproc.h"
#include "npscan.h"
#include "pcr.h"
#include "iter.h"
#include "ls.h"
#include "transfer.h"
#include "assemble.h"
#include "ugdevices.h"
#include "ugstruct.h"
#include "misc.h"
#include "general.h"

#include "uggpu.h"

USING_UG_NAMESPACES

// ---- the device library, bound at run time -------------------------------------------------------------------------
#define UGGPU_FUNCS(X)                                                                                                                          \
  X(uggpu_ctx_create) X(uggpu_ctx_destroy) X(uggpu_last_error) X(uggpu_set_fullrefinelevel) X(uggpu_level_create) X(uggpu_level_set_flags)     \
  X(uggpu_mat_set) X(uggpu_transfer_set) X(uggpu_vec_alloc) X(uggpu_vec_upload) X(uggpu_vec_download) X(uggpu_jac_smooth)                       \
  X(uggpu_restrict) X(uggpu_interpolate_correction) X(uggpu_lmgc_preprocess) X(uggpu_lmgc) X(uggpu_ls_defect) X(uggpu_ls_residuum)              \
  X(uggpu_ls_solve) X(uggpu_cg_solve) X(uggpu_bcgs_solve) X(uggpu_launch_count) X(uggpu_smooth) X(uggpu_gs_preprocess) X(uggpu_transfer_set_mode) \
  X(uggpu_dmatcopy) X(uggpu_l_ilubthdecomp) X(uggpu_assemble) X(uggpu_mat_set_pattern) X(uggpu_mat_get) X(uggpu_level_get_flags) \
  X(uggpu_savedata) X(uggpu_loaddata) X(uggpu_minimize_level) X(uggpu_amg_coarsen_rs) X(uggpu_amg_coarsen_vanek) X(uggpu_mat_nnz)

namespace {
struct Api {
#define X(f) decltype(&::f) f = nullptr;
  UGGPU_FUNCS(X)
#undef X
  void *dl = nullptr;
} api;
std::string g_load_error;
}

int gpuls::LoadDeviceLibrary(const char *path)
{
  if (api.dl) return 0;
  const char *cand[3] = {path, getenv("UGGPU_LIB"), "libuggpu.so"};
  for (const char *p : cand) {
    if (!p || !*p) continue;
    api.dl = dlopen(p, RTLD_NOW | RTLD_LOCAL);
    if (api.dl) break;
    g_load_error = dlerror();
  }
  if (!api.dl) return 1;
#define X(f) api.f = (decltype(api.f))dlsym(api.dl, #f); if (!api.f) { g_load_error = "missing symbol " #f; dlclose(api.dl); api.dl = nullptr; return 2; }
  UGGPU_FUNCS(X)
#undef X
  return 0;
}
const char *gpuls::LastLoadError() { return g_load_error.c_str(); }

// ---- device mirror of one multigrid -------------------------------------------------------------------------------------
namespace {

struct Mirror {
  MULTIGRID *mg = nullptr;
  uggpu_ctx *ctx = nullptr;
  int refs = 0;
  int bs = 0;
  const VECDATA_DESC *xdesc = nullptr;
  std::vector<gpuls::FlatLevel> fl;
  std::vector<char> have_level, have_transfer;
  std::vector<const MATDATA_DESC *> level_A;     // what have_level[] refers to
  std::vector<const VECDATA_DESC *> level_x;
  std::map<const void *, int> handles;
  std::vector<double> buf;
  // UG numbers the algebraic levels an AMG transfer builds below level 0 with negative numbers (BOTTOMLEVEL(mg) < 0,
  // np/procs/amgtransfer.cc:737); the device library's levels start at 0.  Device level = UG level + loff; loff is 0 unless a
  // gputransfer with $amg has built such levels (SetBottom below).  The arrays above are indexed through ix().
  int loff = 0;
  int dl(int level) const { return level + loff; }
  static int ix(int level) { return level + MAXLEVEL; }

  int handle(const void *desc)
  {
    auto it = handles.find(desc);
    if (it != handles.end()) return it->second;
    int h = (int)handles.size() + 1;
    handles[desc] = h;
    return h;
  }
};

std::map<MULTIGRID *, Mirror> g_mirrors;

int dev_fail(const char *where)
{
  UserWriteF("gpuls: %s failed: %s\n", where, api.uggpu_last_error ? api.uggpu_last_error() : "device library not loaded");
  return 1;
}
#define DEV(call) do { if (api.call) return dev_fail(#call); } while (0)

Mirror *Acquire(MULTIGRID *mg)
{
  if (!api.dl && gpuls::LoadDeviceLibrary(NULL)) {
    UserWriteF("gpuls: cannot load the device library (%s); there is no CPU fallback\n", g_load_error.c_str());
    return NULL;
  }
  Mirror &m = g_mirrors[mg];
  if (m.refs == 0) {
    m.mg = mg;
    int dev = 0;
    if (const char *s = getenv("UGGPU_DEVICE")) dev = atoi(s);
    if (api.uggpu_ctx_create(dev, &m.ctx)) { dev_fail("uggpu_ctx_create"); g_mirrors.erase(mg); return NULL; }
    m.fl.assign(2 * MAXLEVEL, gpuls::FlatLevel());
    m.have_level.assign(2 * MAXLEVEL, 0);
    m.have_transfer.assign(2 * MAXLEVEL, 0);
    m.level_A.assign(2 * MAXLEVEL, (const MATDATA_DESC *)NULL);
    m.level_x.assign(2 * MAXLEVEL, (const VECDATA_DESC *)NULL);
    m.handles.clear();
    m.loff = 0;
  }
  m.refs++;
  return &m;
}

void Release(MULTIGRID *mg)
{
  auto it = g_mirrors.find(mg);
  if (it == g_mirrors.end()) return;
  if (--it->second.refs > 0) return;
  if (it->second.ctx) api.uggpu_ctx_destroy(it->second.ctx);
  g_mirrors.erase(it);
}

Mirror *Find(MULTIGRID *mg)
{
  auto it = g_mirrors.find(mg);
  return it == g_mirrors.end() ? NULL : &it->second;
}

// A PreProcess failed somewhere inside a bracket: the numprocs above it return their error without a PostProcess, so their mirror
// references would never be given back and the NEXT solve would find a mirror that still claims to hold the (then reassembled) matrix.
// The mirror is therefore dropped as a whole -- device context, uploaded levels, all references; the next PreProcess starts clean.
void Destroy(MULTIGRID *mg)
{
  auto it = g_mirrors.find(mg);
  if (it == g_mirrors.end()) return;
  if (it->second.ctx) api.uggpu_ctx_destroy(it->second.ctx);
  g_mirrors.erase(it);
}

// the mirror a numproc acquired in its PreProcess, if it still exists (Destroy after a failed PreProcess removes it)
Mirror *Live(Mirror *held, MULTIGRID *mg)
{
  Mirror *m = Find(mg);
  return (held != NULL && m == held) ? m : NULL;
}
#define PRE_FAIL(np_, result_) do { Destroy(NP_MG(theNP)); (np_)->m = NULL; NP_RETURN(1, result_); } while (0)
#define PRE_FAIL_JAC(np_, result_) do { (np_)->acquired = 0; PRE_FAIL(np_, result_); } while (0)

// An AMG transfer has built levels down to `bottom` < 0: they become device levels 0.. and everything above moves up.  Levels that
// were uploaded under another offset are flattened again (at most once per bracket: the AMG runs first in a transfer's PreProcess).
void SetBottom(Mirror *m, int bottom)
{
  const int loff = bottom < 0 ? -bottom : 0;
  if (loff == m->loff) return;
  m->loff = loff;
  std::fill(m->have_level.begin(), m->have_level.end(), 0);
  std::fill(m->have_transfer.begin(), m->have_transfer.end(), 0);
}

// flatten + upload matrix A and the row flags of `level` (once per PreProcess bracket)
int EnsureLevel(Mirror *m, int level, const VECDATA_DESC *x, const MATDATA_DESC *A)
{
  // one upload per PreProcess bracket and (level, A, x): another matrix or vector descriptor on the same level is flattened again
  const int k = Mirror::ix(level);
  if (level < -MAXLEVEL || level >= MAXLEVEL || m->dl(level) < 0 || m->dl(level) >= UGGPU_MAX_LEVELS) { UserWriteF("gpuls: level %d is outside the device library's range\n", level); return 1; }
  if (m->have_level[k] == 2) return 0;        // built on the device (gputransfer $gpuamg): nothing to flatten
  if (m->have_level[k] && m->level_A[k] == A && m->level_x[k] == x) return 0;
  if (m->have_level[k]) { m->have_level[k] = 0; m->have_transfer[k] = 0; if (level + 1 < MAXLEVEL) m->have_transfer[k + 1] = 0; }
  gpuls::FlatLevel &f = m->fl[k];
  if (gpuls::FlattenFlags(m->mg, level, x, f)) { UserWriteF("gpuls: level %d is not a pure nodal vector format\n", level); return 1; }
  if (f.bs > UGGPU_MAX_BS) { UserWriteF("gpuls: %d components per vector exceed UGGPU_MAX_BS\n", f.bs); return 1; }
  if (gpuls::FlattenMatrix(m->mg, level, A, f)) { UserWrite("gpuls: cannot flatten the matrix\n"); return 1; }
  m->bs = f.bs;
  m->xdesc = x;
  DEV(uggpu_level_create(m->ctx, m->dl(level), f.n, f.bs));
  DEV(uggpu_level_set_flags(m->ctx, m->dl(level), f.vclass.data(), f.vnclass.data(), f.ctl.data(), f.skip.data()));
  DEV(uggpu_mat_set(m->ctx, m->dl(level), m->handle(A), f.rowptr.data(), f.col.data(), f.val.data()));
  DEV(uggpu_set_fullrefinelevel(m->ctx, m->dl(FULLREFINELEVEL(m->mg))));
  m->have_level[k] = 1;
  m->level_A[k] = A; m->level_x[k] = x;
  return 0;
}

// imat: `transfer $M` -- the stencils come from the stored interpolation matrices (gpuls_flatten.h FlattenTransferIMAT)
int EnsureTransfer(Mirror *m, int level, int imat)
{
  const int k = Mirror::ix(level);
  if (level < 1) imat = 1;       // below level 1 the reference always works on the stored interpolation matrices (transfer.cc:733, :756)
  if (m->have_transfer[k] == 1 + imat) return 0;
  if (level <= -MAXLEVEL || !m->have_level[k] || !m->have_level[k - 1]) return 1;
  gpuls::FlatLevel &f = m->fl[k];
  if (int rc = imat ? gpuls::FlattenTransferIMAT(m->mg, level, f) : gpuls::FlattenTransfer(m->mg, level, f)) {
    UserWriteF("gpuls: cannot flatten the %s transfer of level %d (code %d)\n", imat ? "IMAT" : "standard", level, rc);
    return 1;
  }
  DEV(uggpu_transfer_set(m->ctx, m->dl(level), f.p_rowptr.data(), f.p_col.data(), f.p_w.data(), f.r_rowptr.data(), f.r_col.data(), f.r_w.data()));
  DEV(uggpu_transfer_set_mode(m->ctx, m->dl(level), imat ? UGGPU_TRANSFER_IMAT : UGGPU_TRANSFER_STANDARD));
  m->have_transfer[k] = 1 + imat;
  return 0;
}

// have_level[] == 2: an algebraic level that exists on the device only (gputransfer $gpuamg): UG holds no vectors for it.  The cycle uses
// such levels as work space (the restriction writes their defect, dset their correction), so "uploading" is allocating and nothing comes back.
int Upload(Mirror *m, int level, const VECDATA_DESC *vd)
{
  if (m->have_level[Mirror::ix(level)] == 2) { DEV(uggpu_vec_alloc(m->ctx, m->dl(level), m->handle(vd))); return 0; }
  const gpuls::FlatLevel &f = m->fl[Mirror::ix(level)];
  m->buf.resize((size_t)f.n * f.bs + 1);
  gpuls::GatherVector(m->mg, level, vd, f.bs, m->buf.data());
  DEV(uggpu_vec_upload(m->ctx, m->dl(level), m->handle(vd), m->buf.data()));
  return 0;
}

int Download(Mirror *m, int level, const VECDATA_DESC *vd)
{
  if (m->have_level[Mirror::ix(level)] == 2) return 0;
  const gpuls::FlatLevel &f = m->fl[Mirror::ix(level)];
  m->buf.resize((size_t)f.n * f.bs + 1);
  DEV(uggpu_vec_download(m->ctx, m->dl(level), m->handle(vd), m->buf.data()));
  gpuls::ScatterVector(m->mg, level, vd, f.bs, m->buf.data());
  return 0;
}

void VsToArray(const VEC_SCALAR vs, int bs, double *out) { for (int i = 0; i < UGGPU_MAX_BS; i++) out[i] = i < bs ? vs[i] : 1.0; }

// =========================================================================================================================
// iter.gpujac  (reference: NP_SMOOTHER iter.cc:152-177, SmootherInit :763, Smoother :817, JacobiPreProcess/Step :894-925)
// iter.gpugs   (reference: class `gs`,  GSPreProcess :1003, GSStep :1039)          -- SURVEY.md 8f.2
// iter.gpusgs  (reference: class `sgs`, SGSPreProcess :1353, SGSSmoother :1392)
// iter.gpusor  (reference: class `sor`, SORPreProcess :4717, SORStep :4744, SORSmoother :4786; $damp is the relaxation omega)
// iter.gpuilu  (reference: class `ilu`, NP_ILU :326-330, ILUInit :5415 ($beta), ILUPreProcess :5444, ILUStep :5478)
// One struct and one set of functions: `kind` (UGGPU_SM_*) is set by the constructor of the class.
// =========================================================================================================================
struct NP_GPUJAC {
  NP_ITER iter;
  VEC_SCALAR damp;
  Mirror *m;
  INT acquired;        // PreProcess is called once per level (iter.cc:7719): one mirror reference each
  INT kind;            // UGGPU_SM_JAC / GS / SGS / SOR / ILU
  int t_handle;        // sgs: the extra temporary NP_SGS_t (iter.cc:1386) lives on the device only
  VEC_SCALAR beta;     // ilu: $beta (NP_ILU.beta iter.cc:328)
  int L_handle;        // ilu: the decomposed copy of A (NP_SMOOTHER.L, AllocMDFromMD iter.cc:5457) lives on the device only
};
static const char *SmootherName(INT kind) { return kind == UGGPU_SM_GS ? "gpugs" : kind == UGGPU_SM_SGS ? "gpusgs" : kind == UGGPU_SM_SOR ? "gpusor" : kind == UGGPU_SM_ILU ? "gpuilu" : "gpujac"; }

INT GpuJacInit(NP_BASE *theNP, INT argc, char **argv)
{
  NP_GPUJAC *np = (NP_GPUJAC *)theNP;
  for (int i = 0; i < MAX_VEC_COMP; i++) np->damp[i] = 1.0;
  sc_read(np->damp, NP_FMT(np), np->iter.b, "damp", argc, argv);          // iter.cc:771
  for (int i = 0; i < MAX_VEC_COMP; i++) np->beta[i] = 0.0;
  if (np->kind == UGGPU_SM_ILU) sc_read(np->beta, NP_FMT(np), np->iter.b, "beta", argc, argv);   // iter.cc:5422-5423
  return NPIterInit(&np->iter, argc, argv);
}

INT GpuJacDisplay(NP_BASE *theNP)
{
  NP_GPUJAC *np = (NP_GPUJAC *)theNP;
  NPIterDisplay(&np->iter);
  UserWrite("configuration parameters:\n");
  if (sc_disp(np->damp, np->iter.b, "damp")) REP_ERR_RETURN(1);
  if (np->kind == UGGPU_SM_ILU && sc_disp(np->beta, np->iter.b, "beta")) REP_ERR_RETURN(1);
  UserWriteF(DISPLAY_NP_FORMAT_SS, "device", "B200 via libuggpu");
  return 0;
}

INT GpuJacPreProcess(NP_ITER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *baselevel, INT *result)
{
  NP_GPUJAC *np = (NP_GPUJAC *)theNP;
  Mirror *m = Acquire(NP_MG(theNP));
  if (m == NULL) NP_RETURN(1, result[0]);
  np->m = m;
  np->acquired++;
  if (EnsureLevel(np->m, level, x, A)) PRE_FAIL_JAC(np, result[0]);
  if (np->kind == UGGPU_SM_ILU) {
    // ILUPreProcess iter.cc:5444-5476: L = copy of A (AllocMDFromMD + dmatcopy), l_ilubthdecomp(L, beta, no threshold)
    double beta[UGGPU_MAX_BS];
    for (int i = 0; i < UGGPU_MAX_BS; i++) beta[i] = i < m->bs ? np->beta[i] : 0.0;
    np->L_handle = m->handle(&np->L_handle);
    if (api.uggpu_dmatcopy(m->ctx, m->dl(level), m->dl(level), UGGPU_ALL_VECTORS, np->L_handle, m->handle(A))) { dev_fail("uggpu_dmatcopy"); PRE_FAIL_JAC(np, result[0]); }
    if (api.uggpu_l_ilubthdecomp(m->ctx, m->dl(level), np->L_handle, beta)) {
      PrintErrorMessage('E', "GpuIluPreProcess", "decomposition failed");      // iter.cc:5470
      { dev_fail("uggpu_l_ilubthdecomp"); PRE_FAIL_JAC(np, result[0]); }
    }
  } else if (np->kind != UGGPU_SM_JAC) {
    // l_setindex (iter.cc:1027): rows are numbered in list order by the flattening; the device builds its level schedule
    if (api.uggpu_gs_preprocess(m->ctx, m->dl(level), m->handle(A))) { dev_fail("uggpu_gs_preprocess"); PRE_FAIL_JAC(np, result[0]); }
    np->t_handle = m->handle(&np->t_handle);
  }
  *baselevel = level;                                                      // iter.cc:908
  return 0;
}

// one damped Jacobi step in defect-correction form: x = damp * Diag(A)^-1 b ; b -= A x   (Smoother iter.cc:817-842)
INT GpuJacIter(NP_ITER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUJAC *np = (NP_GPUJAC *)theNP;
  NPIT_A(theNP) = A; NPIT_c(theNP) = x; NPIT_b(theNP) = b;
  Mirror *m = np->m ? Live(np->m, NP_MG(theNP)) : Find(NP_MG(theNP));
  if (m == NULL || !m->have_level[Mirror::ix(level)]) { UserWriteF("%s: Iter without PreProcess\n", SmootherName(np->kind)); NP_RETURN(1, result[0]); }
  double damp[UGGPU_MAX_BS];
  VsToArray(np->damp, m->bs, damp);
  if (Upload(m, level, b)) NP_RETURN(1, result[0]);
  if (api.uggpu_vec_alloc(m->ctx, m->dl(level), m->handle(x))) NP_RETURN(dev_fail("uggpu_vec_alloc"), result[0]);
  if (np->kind != UGGPU_SM_JAC) {
    if (api.uggpu_smooth(m->ctx, m->dl(level), (int)np->kind, m->handle(x), m->handle(b), m->handle(A), damp, np->kind == UGGPU_SM_ILU ? np->L_handle : np->t_handle)) NP_RETURN(dev_fail("uggpu_smooth"), result[0]);
  } else
  if (api.uggpu_jac_smooth(m->ctx, m->dl(level), m->handle(x), m->handle(b), m->handle(A), damp)) NP_RETURN(dev_fail("uggpu_jac_smooth"), result[0]);
  if (Download(m, level, x) || Download(m, level, b)) NP_RETURN(1, result[0]);
  return 0;
}

INT GpuJacPostProcess(NP_ITER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUJAC *np = (NP_GPUJAC *)theNP;
  if (np->acquired > 0) {
    Release(NP_MG(theNP));
    if (--np->acquired == 0) np->m = NULL;
  }
  return 0;
}

INT GpuSmootherConstruct(NP_BASE *theNP, INT kind)
{
  theNP->Init = GpuJacInit;
  theNP->Display = GpuJacDisplay;
  theNP->Execute = NPIterExecute;
  NP_ITER *np = (NP_ITER *)theNP;
  np->PreProcess = GpuJacPreProcess;
  np->Iter = GpuJacIter;
  np->PostProcess = GpuJacPostProcess;
  ((NP_GPUJAC *)theNP)->kind = kind;
  return 0;
}
INT GpuJacConstruct(NP_BASE *theNP) { return GpuSmootherConstruct(theNP, UGGPU_SM_JAC); }
INT GpuGsConstruct(NP_BASE *theNP) { return GpuSmootherConstruct(theNP, UGGPU_SM_GS); }
INT GpuSgsConstruct(NP_BASE *theNP) { return GpuSmootherConstruct(theNP, UGGPU_SM_SGS); }
INT GpuSorConstruct(NP_BASE *theNP) { return GpuSmootherConstruct(theNP, UGGPU_SM_SOR); }
INT GpuIluConstruct(NP_BASE *theNP) { return GpuSmootherConstruct(theNP, UGGPU_SM_ILU); }

// =========================================================================================================================
// transfer.gputransfer  (reference: NP_STANDARD_TRANSFER transfer.cc:115-133, standard mode, $M and $amg)
// =========================================================================================================================
struct NP_GPUTRANSFER {
  NP_TRANSFER transfer;
  Mirror *m;
  INT fl, tl;
  INT imat;            // $M: IMAT_MODE (transfer.cc:564-572): RestrictByMatrix / InterpolateCorrectionByMatrix
  NP_TRANSFER *amg;    // $amg <numproc>: NP_STANDARD_TRANSFER.amg (transfer.cc:132, :593): a transfer numproc of the host -- the reference's
                       // selectionAMG / clusterAMG (np/procs/amgtransfer.cc) -- whose PreProcess builds algebraic levels below level 0 when
                       // the cycle's base level is <= 0 (transfer.cc:660-664).  The coarsening stays the reference's sequential host code (a setup
                       // step); the levels it leaves (vectors, Galerkin matrices, interpolation matrices) are mirrored like any other
                       // level and the device cycle runs on them with the by-matrix transfer, as the reference does for all levels < 1.
  INT amg_ran;
  INT level;           // $L: level optimisation, AdaptCorrection = MinimizeLevel (transfer.cc:574, :812, :488)
  VECDATA_DESC *t;     // its work vector (transfer.cc:592 $t), on the device only
  INT dirichlet;       // $D [k]: AssembleDirichletBoundary on the levels in PreProcess (transfer.cc:575, :666-678)
  // $gpuamg {RugeStueben | Vanek | VanekPC}: the algebraic levels below level 0 are built by the device library itself (uggpu_amg_coarsen_rs /
  // uggpu_amg_coarsen_vanek: the reference's selectionAMG $strongRel $C RugeStueben $I RugeStueben resp. clusterAMG $strongVanek $C VanekNeuss
  // $I Vanek / PiecewiseConstant, $CM Galerkin, bit for bit) and exist on the device only: UG's grid manager never sees them.  $theta,
  // $vectLimit, $levelLimit as the AMG numprocs' $strong... value, $vectLimit, $levelLimit (amgtransfer.cc:540-552, :800-812).  Needs the
  // device-resident cycle with the device base solver (gpulmgc $devbase): there are no host vectors for a host numproc to work on.
  INT gpuamg;          // 0 none, 1 Ruge-Stueben, 2 Vanek (smoothed aggregation), 3 Vanek with piecewise constant interpolation
  DOUBLE theta;
  INT vectLimit, levelLimit, matLimit;
  DOUBLE bandLimit, vRedLimit, mRedLimit;        // the AMG numprocs' other stopping criteria (amgtransfer.cc:543-550, :806-826, :1000-1012)
};

INT GpuRestrictDefect(NP_TRANSFER *theNP, INT level, VECDATA_DESC *to, VECDATA_DESC *from, MATDATA_DESC *A, VEC_SCALAR damp, INT *result);

INT GpuTransferInit(NP_BASE *theNP, INT argc, char **argv)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  np->imat = ReadArgvOption("M", argc, argv);
  np->amg = (NP_TRANSFER *)ReadArgvNumProc(theNP->mg, "amg", TRANSFER_CLASS_NAME, argc, argv);       // transfer.cc:593
  np->amg_ran = 0;
  if (np->amg != NULL && (np->amg->RestrictDefect == GpuRestrictDefect || np->amg->PreProcess == NULL)) {
    UserWrite("gputransfer: $amg must name a host transfer numproc that builds the algebraic levels (selectionAMG, clusterAMG)\n");
    return NP_NOT_ACTIVE;
  }
  np->level = ReadArgvOption("L", argc, argv);                                                        // transfer.cc:574
  np->t = ReadArgvVecDesc(theNP->mg, "t", argc, argv);                                                // transfer.cc:592
  np->dirichlet = ReadArgvOption("D", argc, argv);                                                    // transfer.cc:575
  np->gpuamg = 0;
  {
    char kind[VALUELEN];
    if (ReadArgvChar("gpuamg", kind, argc, argv) == 0) {
      np->gpuamg = strcmp(kind, "RugeStueben") == 0 ? 1 : strcmp(kind, "Vanek") == 0 ? 2 : strcmp(kind, "VanekPC") == 0 ? 3 : 0;
      if (np->gpuamg == 0 || np->amg != NULL) { UserWrite("gputransfer: $gpuamg {RugeStueben | Vanek | VanekPC}, not together with $amg\n"); return NP_NOT_ACTIVE; }
    }
    np->theta = np->gpuamg == 1 ? 0.25 : 0.08;
    ReadArgvDOUBLE("theta", &np->theta, argc, argv);
    np->vectLimit = 0; ReadArgvINT("vectLimit", &np->vectLimit, argc, argv);
    np->levelLimit = -16; ReadArgvINT("levelLimit", &np->levelLimit, argc, argv);
    np->matLimit = 0; ReadArgvINT("matLimit", &np->matLimit, argc, argv);
    np->bandLimit = 0.0; ReadArgvDOUBLE("bandLimit", &np->bandLimit, argc, argv);
    np->vRedLimit = 0.0; ReadArgvDOUBLE("vRedLimit", &np->vRedLimit, argc, argv);
    np->mRedLimit = 0.0; ReadArgvDOUBLE("mRedLimit", &np->mRedLimit, argc, argv);
    if (np->levelLimit > 0 || np->levelLimit < -MAXLEVEL + 1) { UserWrite("gputransfer: $levelLimit must be in -MAXLEVEL+1..0\n"); return NP_NOT_ACTIVE; }
  }
  if (ReadArgvOption("R", argc, argv) || ReadArgvOption("S", argc, argv)) {
    UserWrite("gputransfer: the standard (geometric) transfer, $M (stored interpolation matrices), $L (level optimisation), $D and $amg are on the GPU path; $R $S are not supported\n");
    return NP_NOT_ACTIVE;
  }
  return NPTransferInit((NP_TRANSFER *)theNP, argc, argv);                // transfer.cc:593
}

INT GpuTransferDisplay(NP_BASE *theNP)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  NPTransferDisplay((NP_TRANSFER *)theNP);
  UserWriteF(DISPLAY_NP_FORMAT_SS, "Restrict", np->imat ? "RestrictByMatrix (device)" : "StandardRestrict (device)");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "InterpolateCor", np->imat ? "InterpolateCorrectionByMatrix (device)" : "StandardInterpolateCorrection (device)");
  if (np->amg != NULL) UserWriteF(DISPLAY_NP_FORMAT_SS, "amg", ENVITEM_NAME(np->amg));
  UserWriteF(DISPLAY_NP_FORMAT_SI, "level", (int)np->level);
  if (np->gpuamg) {
    UserWriteF(DISPLAY_NP_FORMAT_SS, "gpuamg", np->gpuamg == 1 ? "RugeStueben" : np->gpuamg == 2 ? "Vanek" : "VanekPC");
    UserWriteF(DISPLAY_NP_FORMAT_SF, "theta", (float)np->theta);
    UserWriteF(DISPLAY_NP_FORMAT_SI, "vectLimit", (int)np->vectLimit);
    UserWriteF(DISPLAY_NP_FORMAT_SI, "levelLimit", (int)np->levelLimit);
    UserWriteF(DISPLAY_NP_FORMAT_SI, "matLimit", (int)np->matLimit);
    UserWriteF(DISPLAY_NP_FORMAT_SF, "bandLimit", (float)np->bandLimit);
    UserWriteF(DISPLAY_NP_FORMAT_SF, "vRedLimit", (float)np->vRedLimit);
    UserWriteF(DISPLAY_NP_FORMAT_SF, "mRedLimit", (float)np->mRedLimit);
  }
  return 0;
}

// TransferPreProcess transfer.cc:643: nothing to do in the sequential standard mode; here: build the device stencils
INT GpuTransferPreProcess(NP_TRANSFER *theNP, INT *fl, INT tl, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  np->m = Acquire(NP_MG(theNP));
  if (np->m == NULL) NP_RETURN(1, result[0]);
  np->amg_ran = 0;
  if (np->amg != NULL && *fl <= 0) {                                      // transfer.cc:660-664: the AMG sets *fl to the new bottom level
    if ((*np->amg->PreProcess)(np->amg, fl, 0, x, b, A, result)) PRE_FAIL(np, result[0]);
    np->amg_ran = 1;
    SetBottom(np->m, *fl);
  }
  if (np->dirichlet) {                                                    // transfer.cc:666-678: a host step on UG's lists, before they are flattened
    int i = *fl;
    if (np->dirichlet > 1) i = np->dirichlet - 1;
    for (; i <= tl; i++)
      if (AssembleDirichletBoundary(GRID_ON_LEVEL(NP_MG(theNP), i), A, x, b)) PRE_FAIL(np, result[0]);
  }
  if (np->gpuamg && *fl <= 0 && np->levelLimit < 0) {
    // room for the levels the device library may build below level 0, then the coarsening loop of AMGTransferPreProcess (amgtransfer.cc:795-925)
    // with its $vectLimit / $levelLimit criteria; the AMG starts at level 0 like the reference's ("AMG can only be used on levels >= 0", :750)
    Mirror *m = np->m;
    SetBottom(m, np->levelLimit);
    for (int l = -MAXLEVEL; l < 0; l++) { m->have_level[Mirror::ix(l)] = 0; m->have_transfer[Mirror::ix(l)] = 0; }
    m->have_transfer[Mirror::ix(0)] = 0;
    for (int l = 0; l <= tl; l++) if (EnsureLevel(m, l, x, A)) PRE_FAIL(np, result[0]);
    if (m->bs != 1) { UserWrite("gputransfer: $gpuamg handles scalar equations\n"); PRE_FAIL(np, result[0]); }
    // nMat = 2 * nCon of the reference (:802): a diagonal connection counts once, a pair of off-diagonal matrices once -> entries + vectors
    int level = 0, nvec = m->fl[Mirror::ix(0)].n;
    double nmat = (double)m->fl[Mirror::ix(0)].col.size() + nvec;
    while (level > np->levelLimit) {
      if (np->vectLimit != 0 && nvec <= np->vectLimit) break;                                     // :806
      if (np->matLimit != 0 && nmat <= np->matLimit) break;                                       // :813
      if (np->bandLimit != 0.0 && nmat / (double)nvec > np->bandLimit) break;                     // :820
      int nc = 0;
      const int rc = np->gpuamg == 1 ? api.uggpu_amg_coarsen_rs(m->ctx, m->dl(level), m->handle(A), np->theta, &nc)
                                     : api.uggpu_amg_coarsen_vanek(m->ctx, m->dl(level), m->handle(A), np->theta, np->gpuamg == 2 ? 1 : 0, &nc);
      if (rc) { dev_fail("uggpu_amg_coarsen"); PRE_FAIL(np, result[0]); }
      if (nc == 0) break;                                                                        // all or no vectors coarse: the coarsening has come to its end
      m->have_level[Mirror::ix(level - 1)] = 2;
      m->have_transfer[Mirror::ix(level)] = 2;
      gpuls::FlatLevel &f = m->fl[Mirror::ix(level - 1)];
      f = gpuls::FlatLevel(); f.n = nc; f.bs = 1;
      const double cmat = (double)api.uggpu_mat_nnz(m->ctx, m->dl(level - 1), m->handle(A)) + nc;
      // the reduction criteria are tested AFTER the level was built, and the level stays (:1000-1012)
      const bool stalled = (np->vRedLimit != 0.0 && (double)nc / (double)nvec > np->vRedLimit) || (np->mRedLimit != 0.0 && cmat / nmat > np->mRedLimit);
      nvec = nc; nmat = cmat;
      level--;
      if (stalled) break;
    }
    *fl = level;
    np->fl = *fl; np->tl = tl;
    for (int l = 1; l <= tl; l++) if (EnsureTransfer(m, l, np->imat ? 1 : 0)) PRE_FAIL(np, result[0]);
    return 0;
  }
  np->fl = *fl; np->tl = tl;
  bool bad = false;
  for (int l = *fl; l <= tl && !bad; l++) if (EnsureLevel(np->m, l, x, A)) bad = true;
  for (int l = *fl + 1; l <= tl && !bad; l++) if (EnsureTransfer(np->m, l, np->imat ? 1 : 0)) bad = true;
  if (bad) {
    // the AMG numproc's bracket is closed again (it frees the matrix descriptors of its levels and disposes them, amgtransfer.cc:1166-1200):
    // no PostProcess will follow a failed PreProcess
    if (np->amg_ran && np->amg->PostProcess != NULL) { INT r2 = 0; (*np->amg->PostProcess)(np->amg, fl, 0, x, b, A, &r2); np->amg_ran = 0; }
    PRE_FAIL(np, result[0]);
  }
  return 0;
}

// RestrictDefect transfer.cc:724: fine `level` -> level-1
INT GpuRestrictDefect(NP_TRANSFER *theNP, INT level, VECDATA_DESC *to, VECDATA_DESC *from, MATDATA_DESC *A, VEC_SCALAR damp, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  Mirror *m = np->m ? Live(np->m, NP_MG(theNP)) : Find(NP_MG(theNP));
  if (m == NULL || level <= -MAXLEVEL || level >= MAXLEVEL || !m->have_transfer[Mirror::ix(level)]) { UserWrite("gputransfer: RestrictDefect without PreProcess\n"); NP_RETURN(1, result[0]); }
  double d[UGGPU_MAX_BS];
  VsToArray(damp, m->bs, d);
  // the coarse vector is an input too: rows with VNCLASS < NEWDEF_CLASS keep their values (transgrid.cc:143-147)
  if (Upload(m, level, from) || Upload(m, level - 1, to)) NP_RETURN(1, result[0]);
  if (api.uggpu_restrict(m->ctx, m->dl(level), m->handle(to), m->handle(from), d)) NP_RETURN(dev_fail("uggpu_restrict"), result[0]);
  if (Download(m, level - 1, to)) NP_RETURN(1, result[0]);
  result[0] = 0;
  return 0;
}

// InterpolateCorrection transfer.cc:747: level-1 -> fine `level`
INT GpuInterpolateCorrection(NP_TRANSFER *theNP, INT level, VECDATA_DESC *to, VECDATA_DESC *from, MATDATA_DESC *A, VEC_SCALAR damp, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  Mirror *m = np->m ? Live(np->m, NP_MG(theNP)) : Find(NP_MG(theNP));
  if (m == NULL || level <= -MAXLEVEL || level >= MAXLEVEL || !m->have_transfer[Mirror::ix(level)]) { UserWrite("gputransfer: InterpolateCorrection without PreProcess\n"); NP_RETURN(1, result[0]); }
  double d[UGGPU_MAX_BS];
  VsToArray(damp, m->bs, d);
  if (Upload(m, level - 1, from)) NP_RETURN(1, result[0]);
  if (api.uggpu_vec_alloc(m->ctx, m->dl(level), m->handle(to))) NP_RETURN(dev_fail("uggpu_vec_alloc"), result[0]);
  if (api.uggpu_interpolate_correction(m->ctx, m->dl(level), m->handle(to), m->handle(from), d)) NP_RETURN(dev_fail("uggpu_interpolate_correction"), result[0]);
  if (Download(m, level, to)) NP_RETURN(1, result[0]);
  result[0] = 0;
  return 0;
}

// AdaptCorrection transfer.cc:812 (called by Lmgc after the post-smoothing, iter.cc:7944): with $L, MinimizeLevel (:488) on c and b of `level`
INT GpuAdaptCorrection(NP_TRANSFER *theNP, INT level, VECDATA_DESC *c, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  if (!np->level) return 0;
  Mirror *m = np->m ? Live(np->m, NP_MG(theNP)) : Find(NP_MG(theNP));
  if (m == NULL || level <= -MAXLEVEL || level >= MAXLEVEL || !m->have_level[Mirror::ix(level)]) { UserWrite("gputransfer: AdaptCorrection without PreProcess\n"); NP_RETURN(1, result[0]); }
  if (Upload(m, level, c) || Upload(m, level, b)) NP_RETURN(1, result[0]);
  if (api.uggpu_minimize_level(m->ctx, m->dl(level), m->handle(c), m->handle(b), m->handle(A), m->handle(&np->t))) NP_RETURN(dev_fail("uggpu_minimize_level"), result[0]);
  if (Download(m, level, c) || Download(m, level, b)) NP_RETURN(1, result[0]);
  return 0;
}

// InterpolateNewVectors / ProjectSolution (transfer.cc:772, :792): hooks of nested iteration (nonlinear solvers, time steppers), called once per
// grid adaption and not inside the cycle.  They act on UG's VVALUEs on the host with the reference's own functions; every entry point of the
// gpuls classes uploads its vector arguments on entry, so no device copy can go stale.
INT GpuInterpolateNewVectors(NP_TRANSFER *theNP, INT fl, INT tl, VECDATA_DESC *x, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  for (INT i = fl + 1; i <= tl; i++) {
    result[0] = np->imat ? InterpolateNewVectorsByMatrix(GRID_ON_LEVEL(NP_MG(theNP), i), x) : StandardInterpolateNewVectors(GRID_ON_LEVEL(NP_MG(theNP), i), x);
    if (result[0]) NP_RETURN(1, result[0]);
  }
  return 0;
}

INT GpuProjectSolution(NP_TRANSFER *theNP, INT fl, INT tl, VECDATA_DESC *x, INT *result)
{
  result[0] = 0;
  for (INT i = tl - 1; i >= fl; i--) {
    result[0] = StandardProject(GRID_ON_LEVEL(NP_MG(theNP), i), x, x);
    if (result[0]) NP_RETURN(1, result[0]);
  }
  return 0;
}

INT GpuTransferPostProcess(NP_TRANSFER *theNP, INT *fl, INT tl, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUTRANSFER *np = (NP_GPUTRANSFER *)theNP;
  INT rc = 0;
  if (np->amg != NULL && np->amg_ran && np->amg->PostProcess != NULL) {      // transfer.cc:836-838: disposes the algebraic levels (unless $hold)
    if ((*np->amg->PostProcess)(np->amg, fl, 0, x, b, A, result)) rc = 1;
    np->amg_ran = 0;
    // the levels below 0 are gone from the host (or will be rebuilt by the next PreProcess): nothing uploaded for them may be reused
    if (Mirror *m = Live(np->m, NP_MG(theNP)))
      for (int l = -MAXLEVEL; l <= 0; l++) { m->have_level[Mirror::ix(l)] = 0; m->have_transfer[Mirror::ix(l)] = 0; if (l == 0) m->have_transfer[Mirror::ix(1)] = 0; }
  }
  if (np->m) Release(NP_MG(theNP));
  np->m = NULL;
  if (rc) REP_ERR_RETURN(1);
  return 0;
}

INT GpuTransferConstruct(NP_BASE *theNP)
{
  theNP->Init = GpuTransferInit;
  theNP->Display = GpuTransferDisplay;
  theNP->Execute = NPTransferExecute;
  NP_TRANSFER *np = (NP_TRANSFER *)theNP;
  np->PreProcess = GpuTransferPreProcess;
  np->PreProcessProject = NULL;
  np->PreProcessSolution = NULL;
  np->InterpolateCorrection = GpuInterpolateCorrection;
  np->RestrictDefect = GpuRestrictDefect;
  np->InterpolateNewVectors = GpuInterpolateNewVectors;      // nested-iteration hooks: the reference's functions on the host
  np->ProjectSolution = GpuProjectSolution;
  np->AdaptCorrection = GpuAdaptCorrection;   // does nothing without $L, like AdaptCorrection transfer.cc:812
  np->PostProcess = GpuTransferPostProcess;
  np->PostProcessProject = NULL;
  np->PostProcessSolution = NULL;
  return 0;
}

// =========================================================================================================================
// iter.gpulmgc  (reference: NP_LMGC iter.cc:411-429, LmgcInit :7613, LmgcPreProcess :7707, Lmgc :7741, LmgcPostProcess :7951)
// =========================================================================================================================
struct NP_GPULMGC {
  NP_ITER iter;
  INT gamma, nu1, nu2, baselevel;
  NP_TRANSFER *Transfer;
  NP_ITER *PreSmooth, *PostSmooth;
  NP_LINEAR_SOLVER *BaseSolver;
  VECDATA_DESC *t;
  VEC_SCALAR damp;
  INT devbase, unfused;
  Mirror *m;
  INT level;                    // level of the PreProcess bracket
  // descriptors of the cycle in flight (for the host base-solver callback)
  VECDATA_DESC *cur_c, *cur_b;
  MATDATA_DESC *cur_A;
  int t_handle;
};

INT GpuLmgcInit(NP_BASE *theNP, INT argc, char **argv)
{
  NP_GPULMGC *np = (NP_GPULMGC *)theNP;
  char post[VALUELEN], pre[VALUELEN], base[VALUELEN];
  np->t = ReadArgvVecDesc(theNP->mg, "t", argc, argv);
  np->Transfer = (NP_TRANSFER *)ReadArgvNumProc(theNP->mg, "T", TRANSFER_CLASS_NAME, argc, argv);
  for (int i = 1; i < argc; i++)
    if (argv[i][0] == 'S') {
      if (sscanf(argv[i], "S %s %s %s", pre, post, base) != 3) continue;
      np->PreSmooth = (NP_ITER *)GetNumProcByName(theNP->mg, pre, ITER_CLASS_NAME);
      np->PostSmooth = (NP_ITER *)GetNumProcByName(theNP->mg, post, ITER_CLASS_NAME);
      np->BaseSolver = (NP_LINEAR_SOLVER *)GetNumProcByName(theNP->mg, base, LINEAR_SOLVER_CLASS_NAME);
      break;
    }
  if (ReadArgvINT("g", &(np->gamma), argc, argv)) np->gamma = 1;
  if (ReadArgvINT("n1", &(np->nu1), argc, argv)) np->nu1 = 1;
  if (ReadArgvINT("n2", &(np->nu2), argc, argv)) np->nu2 = 1;
  if (ReadArgvINT("b", &(np->baselevel), argc, argv)) np->baselevel = 0;
  if (np->baselevel < 0) {                                                // iter.cc:7645-7651
    int i;
    for (i = FULLREFINELEVEL(NP_MG(theNP)); i > 0; i--)
      if (NVEC(GRID_ON_LEVEL(NP_MG(theNP), i)) <= -np->baselevel) break;
    np->baselevel = i;
  }
  np->devbase = ReadArgvOption("devbase", argc, argv);
  np->unfused = ReadArgvOption("unfused", argc, argv);
  if (np->Transfer == NULL || np->PreSmooth == NULL || np->PostSmooth == NULL) REP_ERR_RETURN(NP_NOT_ACTIVE);
  if (np->BaseSolver == NULL && !np->devbase) REP_ERR_RETURN(NP_NOT_ACTIVE);
  if (np->PreSmooth->Iter != GpuJacIter || np->PostSmooth->Iter != GpuJacIter) {
    UserWrite("gpulmgc: $S pre and post smoother must be of class gpujac, gpugs, gpusgs, gpusor or gpuilu\n");
    return NP_NOT_ACTIVE;
  }
  if (((NP_GPUJAC *)np->PreSmooth)->kind != ((NP_GPUJAC *)np->PostSmooth)->kind) {
    UserWrite("gpulmgc: pre and post smoother must be of the same class\n");
    return NP_NOT_ACTIVE;
  }
  if (np->Transfer->RestrictDefect != GpuRestrictDefect) {
    UserWrite("gpulmgc: $T must be of class gputransfer\n");
    return NP_NOT_ACTIVE;
  }
  if (np->gamma < 1) { UserWrite("gpulmgc: $g must be >= 1\n"); return NP_NOT_ACTIVE; }
  INT ret = NPIterInit(&np->iter, argc, argv);
  if (sc_read(np->damp, NP_FMT(np), np->iter.b, "damp", argc, argv))
    for (int i = 0; i < MAX_VEC_COMP; i++) np->damp[i] = 1.0;
  return ret;
}

INT GpuLmgcDisplay(NP_BASE *theNP)
{
  NP_GPULMGC *np = (NP_GPULMGC *)theNP;
  NPIterDisplay(&np->iter);
  UserWrite("configuration parameters:\n");
  UserWriteF(DISPLAY_NP_FORMAT_SI, "g", (int)np->gamma);
  UserWriteF(DISPLAY_NP_FORMAT_SI, "n1", (int)np->nu1);
  UserWriteF(DISPLAY_NP_FORMAT_SI, "n2", (int)np->nu2);
  UserWriteF(DISPLAY_NP_FORMAT_SI, "baselevel", (int)np->baselevel);
  UserWriteF(DISPLAY_NP_FORMAT_SS, "T", np->Transfer ? ENVITEM_NAME(np->Transfer) : "---");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "pre", np->PreSmooth ? ENVITEM_NAME(np->PreSmooth) : "---");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "post", np->PostSmooth ? ENVITEM_NAME(np->PostSmooth) : "---");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "base", np->devbase ? "device LU" : (np->BaseSolver ? ENVITEM_NAME(np->BaseSolver) : "---"));
  UserWriteF(DISPLAY_NP_FORMAT_SS, "schedule", np->unfused ? "one kernel per call" : "fused");
  if (sc_disp(np->damp, np->iter.b, "damp")) REP_ERR_RETURN(1);
  return 0;
}

// base level: exactly what Lmgc does there (iter.cc:7760-7800), on the host numproc, between a download and an upload
int HostBaseSolver(void *user, uggpu_ctx *ctx, int level, int c, int b, int A)
{
  NP_GPULMGC *np = (NP_GPULMGC *)user;
  Mirror *m = np->m;
  LRESULT lresult;
  // the cycle may run on other vectors than the ones the solver was called with (bcgs: Iter(q, p) and Iter(q, s), ls.cc:1944,1990):
  // map the handles back to their descriptors
  VECDATA_DESC *cd = np->cur_c, *bd = np->cur_b;
  for (auto &kv : m->handles) {
    if (kv.second == c) cd = (VECDATA_DESC *)kv.first;
    if (kv.second == b) bd = (VECDATA_DESC *)kv.first;
  }
  level -= m->loff;        // the device library's level -> UG's
  if (Download(m, level, cd) || Download(m, level, bd)) return 1;
  if ((*np->BaseSolver->Residuum)(np->BaseSolver, MIN(level, np->baselevel), level, cd, bd, np->cur_A, &lresult)) return 1;
  if ((*np->BaseSolver->Solver)(np->BaseSolver, level, cd, bd, np->cur_A, np->BaseSolver->abslimit, np->BaseSolver->reduction, &lresult)) return 1;
  if (Upload(m, level, cd) || Upload(m, level, bd)) return 1;
  return 0;
}

void FillCfg(NP_GPULMGC *np, uggpu_lmgc_cfg *cfg)
{
  Mirror *m = np->m;
  memset(cfg, 0, sizeof *cfg);
  cfg->nu1 = np->nu1; cfg->nu2 = np->nu2; cfg->gamma = np->gamma; cfg->baselevel = m->dl(np->baselevel);
  VsToArray(((NP_GPUJAC *)np->PreSmooth)->damp, m->bs, cfg->smooth_damp);
  VsToArray(np->damp, m->bs, cfg->cycle_damp);
  cfg->t = np->t_handle;
  cfg->fused = np->unfused ? 0 : 1;
  cfg->smoother = (int)((NP_GPUJAC *)np->PreSmooth)->kind;
  cfg->level_opt = ((NP_GPUTRANSFER *)np->Transfer)->level ? 1 : 0;          // iter.cc:7944: the transfer's AdaptCorrection inside the cycle
  if (cfg->smoother == UGGPU_SM_ILU) {
    cfg->smoother_L = ((NP_GPUJAC *)np->PreSmooth)->L_handle;
    for (int i = 0; i < UGGPU_MAX_BS; i++) cfg->ilu_beta[i] = i < m->bs ? ((NP_GPUJAC *)np->PreSmooth)->beta[i] : 0.0;
  }
  if (np->devbase) {
    cfg->base_solver = NULL;
    // the parameters of the reference's `ls $I lu` base solver if one is given, else its documented defaults
    cfg->base_maxit = 10; cfg->base_reduction = 1e-8; cfg->base_abslimit = 1e-10;
    if (np->BaseSolver) { cfg->base_reduction = np->BaseSolver->reduction[0]; cfg->base_abslimit = np->BaseSolver->abslimit[0]; }
  } else {
    cfg->base_solver = HostBaseSolver;
    cfg->base_user = np;
  }
}

INT GpuLmgcPreProcess(NP_ITER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *baselevel, INT *result)
{
  NP_GPULMGC *np = (NP_GPULMGC *)theNP;
  np->m = Acquire(NP_MG(theNP));
  if (np->m == NULL) NP_RETURN(1, result[0]);
  np->level = level;
  // same order as LmgcPreProcess iter.cc:7714-7738
  if ((*np->Transfer->PreProcess)(np->Transfer, &(np->baselevel), level, x, b, A, result)) PRE_FAIL(np, result[0]);
  for (int i = np->baselevel + 1; i <= level; i++)
    if ((*np->PreSmooth->PreProcess)(np->PreSmooth, i, x, b, A, baselevel, result)) PRE_FAIL(np, result[0]);
  if (np->PreSmooth != np->PostSmooth)
    for (int i = np->baselevel + 1; i <= level; i++)
      if ((*np->PostSmooth->PreProcess)(np->PostSmooth, i, x, b, A, baselevel, result)) PRE_FAIL(np, result[0]);
  *baselevel = MIN(np->baselevel, level);
  if (!np->devbase && *baselevel >= -MAXLEVEL && np->m->have_level[Mirror::ix(*baselevel)] == 2) {
    UserWrite("gpulmgc: the base level was built on the device (gputransfer $gpuamg): a host base solver has no vectors there, use $devbase\n");
    PRE_FAIL(np, result[0]);
  }
  if (!np->devbase && np->BaseSolver->PreProcess != NULL)
    if ((*np->BaseSolver->PreProcess)(np->BaseSolver, *baselevel, x, b, A, baselevel, result)) PRE_FAIL(np, result[0]);
  if (((NP_GPUJAC *)np->PreSmooth)->damp[0] != ((NP_GPUJAC *)np->PostSmooth)->damp[0]) {
    UserWrite("gpulmgc: pre and post smoother must use the same damping\n");
    PRE_FAIL(np, result[0]);
  }
  np->t_handle = np->m->handle(&np->t);      // the temporary np->t of Lmgc (iter.cc:7810) lives on the device only
  uggpu_lmgc_cfg cfg;
  FillCfg(np, &cfg);
  if (api.uggpu_lmgc_preprocess(np->m->ctx, &cfg, np->m->dl(level), np->m->handle(A))) { dev_fail("uggpu_lmgc_preprocess"); PRE_FAIL(np, result[0]); }
  return 0;
}

// Lmgc iter.cc:7741: c (in/out) and b (in/out) on `level`; the levels below are work space whose final contents the
// reference leaves in the VECTORs, so they are downloaded too.
INT GpuLmgcIter(NP_ITER *theNP, INT level, VECDATA_DESC *c, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPULMGC *np = (NP_GPULMGC *)theNP;
  NPIT_A(theNP) = A; NPIT_c(theNP) = c; NPIT_b(theNP) = b;
  Mirror *m = Live(np->m, NP_MG(theNP));
  if (m == NULL) { UserWrite("gpulmgc: Iter without PreProcess\n"); NP_RETURN(1, result[0]); }
  np->cur_c = c; np->cur_b = b; np->cur_A = A;
  uggpu_lmgc_cfg cfg;
  FillCfg(np, &cfg);
  const int bl = MIN(np->baselevel, level);
  for (int l = bl; l <= level; l++)
    if (Upload(m, l, c) || Upload(m, l, b)) NP_RETURN(1, result[0]);
  if (api.uggpu_lmgc(m->ctx, &cfg, m->dl(level), m->handle(c), m->handle(b), m->handle(A))) NP_RETURN(dev_fail("uggpu_lmgc"), result[0]);
  for (int l = bl; l <= level; l++)
    if (Download(m, l, c) || Download(m, l, b)) NP_RETURN(1, result[0]);
  return 0;
}

INT GpuLmgcPostProcess(NP_ITER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPULMGC *np = (NP_GPULMGC *)theNP;
  // reverse order of LmgcPostProcess iter.cc:7951-7984
  if (!np->devbase && np->BaseSolver->PostProcess != NULL)
    if ((*np->BaseSolver->PostProcess)(np->BaseSolver, np->baselevel, x, b, A, result)) REP_ERR_RETURN(1);
  if (np->PreSmooth != np->PostSmooth)
    for (int i = level; i >= np->baselevel + 1; i--)
      if ((*np->PostSmooth->PostProcess)(np->PostSmooth, i, x, b, A, result)) REP_ERR_RETURN(1);
  for (int i = level; i >= np->baselevel + 1; i--)
    if ((*np->PreSmooth->PostProcess)(np->PreSmooth, i, x, b, A, result)) REP_ERR_RETURN(1);
  if ((*np->Transfer->PostProcess)(np->Transfer, &(np->baselevel), level, x, b, A, result)) REP_ERR_RETURN(1);
  if (np->m) Release(NP_MG(theNP));
  np->m = NULL;
  return 0;
}

INT GpuLmgcConstruct(NP_BASE *theNP)
{
  theNP->Init = GpuLmgcInit;
  theNP->Display = GpuLmgcDisplay;
  theNP->Execute = NPIterExecute;
  NP_ITER *np = (NP_ITER *)theNP;
  np->PreProcess = GpuLmgcPreProcess;
  np->Iter = GpuLmgcIter;
  np->PostProcess = GpuLmgcPostProcess;
  return 0;
}

// =========================================================================================================================
// linear_solver.gpuls  (reference: NP_LS ls.cc:79-108, LinearSolverInit :771, PreProcess :539, Defect :562, Residuum :577,
//                       LinearSolver :637-749 with Update = LSUpdate :869)
// =========================================================================================================================
// linear_solver.gpucg    (reference: class `cg`, NP_CG ls.cc:111-124, CGInit :939, CGPrepare :976, CGUpdate :989, CGClose :1159)
// linear_solver.gpubcgs  (reference: class `bcgs`, NP_BCGS ls.cc:165-187, BCGSInit :1750, BCGSPreProcess :1805, BCGSSolver :1864)
// share everything with gpuls except the iteration that runs on the device.
enum { GPULS_LS = 0, GPULS_CG = 1, GPULS_BCGS = 2 };
struct NP_GPULS {
  NP_LINEAR_SOLVER ls;
  NP_ITER *Iter;
  INT maxiter, baselevel, display;
  VECDATA_DESC *c;
  Mirror *m;
  INT kind;                       // GPULS_*
  INT restart;                    // bcgs $R
  VEC_SCALAR weight;              // bcgs $weight (as given; squared on the device like BCGSInit)
  VECDATA_DESC *w[6];             // cg: p t   bcgs: r p v s t q
};
const char *KindName(INT kind) { return kind == GPULS_CG ? "gpucg" : (kind == GPULS_BCGS ? "gpubcgs" : "gpuls"); }

INT GpuLsInit(NP_BASE *theNP, INT argc, char **argv)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  if (ReadArgvINT("m", &(np->maxiter), argc, argv)) REP_ERR_RETURN(NP_NOT_ACTIVE);
  np->display = ReadArgvDisplay(argc, argv);
  np->Iter = (NP_ITER *)ReadArgvNumProc(theNP->mg, "I", ITER_CLASS_NAME, argc, argv);
  if (np->Iter == NULL) REP_ERR_RETURN(NP_NOT_ACTIVE);
  if (np->Iter->Iter != GpuLmgcIter) {
    UserWriteF("%s: $I must be of class gpulmgc (the device-resident solve has no host iteration)\n", KindName(np->kind));
    return NP_NOT_ACTIVE;
  }
  np->baselevel = 0;
  np->c = ReadArgvVecDesc(theNP->mg, "c", argc, argv);
  for (int i = 0; i < 6; i++) np->w[i] = NULL;
  if (np->kind == GPULS_CG) {
    np->w[0] = ReadArgvVecDesc(theNP->mg, "p", argc, argv);
    np->w[1] = ReadArgvVecDesc(theNP->mg, "t", argc, argv);
  }
  if (np->kind == GPULS_BCGS) {
    const char *nm[6] = {"r", "p", "v", "s", "t", "q"};
    for (int i = 0; i < 6; i++) np->w[i] = ReadArgvVecDesc(theNP->mg, nm[i], argc, argv);
    if (ReadArgvINT("R", &(np->restart), argc, argv)) np->restart = 0;
    if (np->restart < 0) REP_ERR_RETURN(NP_NOT_ACTIVE);
    INT rc = NPLinearSolverInit(&np->ls, argc, argv);       // needs the format of x for sc_read
    if (sc_read(np->weight, NP_FMT(np), NULL, "weight", argc, argv))
      for (int i = 0; i < MAX_VEC_COMP; i++) np->weight[i] = 1.0;
    return rc;
  }
  return NPLinearSolverInit(&np->ls, argc, argv);
}

INT GpuLsDisplay(NP_BASE *theNP)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  NPLinearSolverDisplay(&np->ls);
  UserWriteF(DISPLAY_NP_FORMAT_SI, "m", (int)np->maxiter);
  UserWriteF(DISPLAY_NP_FORMAT_SI, "baselevel", (int)np->baselevel);
  UserWriteF(DISPLAY_NP_FORMAT_SS, "Iter", np->Iter ? ENVITEM_NAME(np->Iter) : "---");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "DispMode", np->display == PCR_NO_DISPLAY ? "NO_DISPLAY" : (np->display == PCR_RED_DISPLAY ? "RED_DISPLAY" : "FULL_DISPLAY"));
  return 0;
}

INT GpuLsPreProcess(NP_LINEAR_SOLVER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *baselevel, INT *result)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  NPLS_A(theNP) = A; NPLS_x(theNP) = x; NPLS_b(theNP) = b;
  np->m = Acquire(NP_MG(theNP));
  if (np->m == NULL) NP_RETURN(1, result[0]);
  if ((*np->Iter->PreProcess)(np->Iter, level, x, b, A, baselevel, result)) PRE_FAIL(np, result[0]);
  np->baselevel = MIN(*baselevel, level);
  // ON_SURFACE loops touch levels FULLREFINELEVEL..level
  for (int l = MIN(np->baselevel, (INT)FULLREFINELEVEL(NP_MG(theNP))); l <= level; l++)
    if (EnsureLevel(np->m, l, x, A)) PRE_FAIL(np, result[0]);
  return 0;
}

INT GpuLsDefect(NP_LINEAR_SOLVER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  Mirror *m = Live(np->m, NP_MG(theNP));
  if (m == NULL) { UserWrite("gpuls: Defect without PreProcess\n"); NP_RETURN(1, result[0]); }
  const int bl = MIN(FULLREFINELEVEL(NP_MG(theNP)), MAX(0, np->baselevel));   // ls.cc:570
  const int fr = MIN((INT)FULLREFINELEVEL(NP_MG(theNP)), level);
  for (int l = fr; l <= level; l++)
    if (Upload(m, l, x) || Upload(m, l, b)) NP_RETURN(1, result[0]);
  if (api.uggpu_ls_defect(m->ctx, m->dl(bl), m->dl(level), m->handle(x), m->handle(b), m->handle(A))) NP_RETURN(dev_fail("uggpu_ls_defect"), result[0]);
  for (int l = fr; l <= level; l++)
    if (Download(m, l, b)) NP_RETURN(1, result[0]);
  return *result;
}

INT GpuLsResiduum(NP_LINEAR_SOLVER *theNP, INT bl, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, LRESULT *lresult)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  Mirror *m = Live(np->m, NP_MG(theNP));
  if (m == NULL) { UserWrite("gpuls: Residuum without PreProcess\n"); NP_RETURN(1, lresult->error_code); }
  const int fr = MIN((INT)FULLREFINELEVEL(NP_MG(theNP)), level);
  for (int l = fr; l <= level; l++)
    if (Upload(m, l, b)) NP_RETURN(1, lresult->error_code);
  uggpu_lresult r;
  memset(&r, 0, sizeof r);
  if (api.uggpu_ls_residuum(m->ctx, m->dl(bl), m->dl(level), m->handle(b), &r)) NP_RETURN(dev_fail("uggpu_ls_residuum"), lresult->error_code);
  for (int i = 0; i < m->bs; i++) lresult->last_defect[i] = r.last_defect[i];
  return 0;
}

// LinearSolver ls.cc:637-749, all iterations on the device
INT GpuLsSolver(NP_LINEAR_SOLVER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, VEC_SCALAR abslimit, VEC_SCALAR reduction, LRESULT *lresult)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  NP_GPULMGC *mgc = (NP_GPULMGC *)np->Iter;
  Mirror *m = Live(np->m, NP_MG(theNP));
  INT PrintID = 0;
  char text[DISPLAY_WIDTH + 4];
  if (m == NULL || mgc->m == NULL) { UserWrite("gpuls: Solver without PreProcess\n"); NP_RETURN(1, lresult->error_code); }
  const int bs = m->bs, bl = np->baselevel;
  for (int i = 0; i < VD_NCOMP(x); i++) { NPLS_red(theNP)[i] = reduction[i]; NPLS_abs(theNP)[i] = abslimit[i]; }
  // UG's descriptors can only be allocated on levels its grid manager holds: algebraic levels that exist on the device only (gputransfer
  // $gpuamg) have no VECTORs -- their work vectors are allocated by Upload on the device
  const int hbl = MAX(bl, (int)BOTTOMLEVEL(NP_MG(theNP)));
  if (AllocVDFromVD(NP_MG(theNP), hbl, level, x, &np->c)) NP_RETURN(1, lresult->error_code);   // ls.cc:662
  CenterInPattern(text, DISPLAY_WIDTH, ENVITEM_NAME(np), '*', "\n");
  if (np->display > PCR_NO_DISPLAY)
    if (PreparePCR(x, np->display, text, &PrintID)) NP_RETURN(1, lresult->error_code);
  clock_t clock_start = clock();

  // up: x and b (all cycle levels: the reference's x += c and the cycle's work vectors live on them), c = np->c
  for (int l = bl; l <= level; l++) {
    if (Upload(m, l, x) || Upload(m, l, b)) NP_RETURN(1, lresult->error_code);
    if (api.uggpu_vec_alloc(m->ctx, m->dl(l), m->handle(np->c))) NP_RETURN(dev_fail("uggpu_vec_alloc"), lresult->error_code);
  }
  mgc->cur_c = np->c; mgc->cur_b = b; mgc->cur_A = A;
  uggpu_lmgc_cfg cfg;
  FillCfg(mgc, &cfg);
  uggpu_lresult r;
  memset(&r, 0, sizeof r);
  double absl[UGGPU_MAX_BS], red[UGGPU_MAX_BS];
  for (int i = 0; i < UGGPU_MAX_BS; i++) { absl[i] = abslimit[i < bs ? i : 0]; red[i] = reduction[i < bs ? i : 0]; r.last_defect[i] = i < bs ? lresult->last_defect[i] : 0.0; }
  std::vector<double> history((size_t)MAX(np->maxiter, 1) * bs, 0.0);
  const int nwork = np->kind == GPULS_CG ? 2 : (np->kind == GPULS_BCGS ? 6 : 0);
  for (int i = 0; i < nwork; i++)          // CGPrepare / CGUpdate / BCGSPreProcess allocate these from the same pool
    if (AllocVDFromVD(NP_MG(theNP), hbl, level, x, &np->w[i])) NP_RETURN(1, lresult->error_code);
  if (np->kind == GPULS_LS) {
    if (api.uggpu_ls_solve(m->ctx, &cfg, m->dl(bl), m->dl(level), m->handle(x), m->handle(b), m->handle(A), m->handle(np->c), np->maxiter, absl, red, &r, history.data()))
      NP_RETURN(dev_fail("uggpu_ls_solve"), lresult->error_code);
  } else if (np->kind == GPULS_CG) {
    if (api.uggpu_cg_solve(m->ctx, &cfg, m->dl(bl), m->dl(level), m->handle(x), m->handle(b), m->handle(A), m->handle(np->c), m->handle(np->w[0]), m->handle(np->w[1]),
                           np->maxiter, absl, red, &r, history.data()))
      NP_RETURN(dev_fail("uggpu_cg_solve"), lresult->error_code);
  } else {
    int wh[6];
    double wgt[UGGPU_MAX_BS];
    for (int i = 0; i < 6; i++) wh[i] = m->handle(np->w[i]);
    for (int i = 0; i < UGGPU_MAX_BS; i++) wgt[i] = i < bs ? np->weight[i] : 1.0;
    if (api.uggpu_bcgs_solve(m->ctx, &cfg, m->dl(bl), m->dl(level), m->handle(x), m->handle(b), m->handle(A), wh, wgt, np->restart, np->maxiter, absl, red, &r, history.data()))
      NP_RETURN(dev_fail("uggpu_bcgs_solve"), lresult->error_code);
  }
  for (int i = 0; i < nwork; i++)
    if (FreeVD(NP_MG(theNP), hbl, level, np->w[i])) REP_ERR_RETURN(1);
  // down: x, b, c on every cycle level (what the CPU classes leave in the VECTORs)
  for (int l = bl; l <= level; l++)
    if (Download(m, l, x) || Download(m, l, b) || Download(m, l, np->c)) NP_RETURN(1, lresult->error_code);

  for (int i = 0; i < bs; i++) { lresult->first_defect[i] = r.first_defect[i]; lresult->last_defect[i] = r.last_defect[i]; }
  lresult->converged = r.converged;
  lresult->number_of_linear_iterations = r.number_of_linear_iterations;
  lresult->error_code = r.error_code;
  double ti = (double)(clock() - clock_start) / CLOCKS_PER_SEC;
  if (np->display > PCR_NO_DISPLAY) {
    VEC_SCALAR d;
    for (int i = 0; i < MAX_VEC_COMP; i++) d[i] = 0.0;
    for (int i = 0; i < bs; i++) d[i] = r.first_defect[i];
    if (DoPCR(PrintID, d, PCR_CRATE_SD)) NP_RETURN(1, lresult->error_code);
    const int nhist = np->kind == GPULS_BCGS ? (r.number_of_linear_iterations + 1) / 2 : r.number_of_linear_iterations;
    for (int it = 0; it < nhist; it++) {
      for (int i = 0; i < bs; i++) d[i] = history[(size_t)it * bs + i];
      if (DoPCR(PrintID, d, PCR_CRATE_SD)) NP_RETURN(1, lresult->error_code);
    }
    if (DoPCR(PrintID, lresult->last_defect, PCR_AVERAGE)) NP_RETURN(1, lresult->error_code);
    if (PostPCR(PrintID, ":ls:avg")) NP_RETURN(1, lresult->error_code);
    if (SetStringValue(":ls:avg:iter", (DOUBLE)(r.number_of_linear_iterations + (r.converged ? 0 : 1)))) NP_RETURN(1, lresult->error_code);
    if (lresult->number_of_linear_iterations != 0)
      UserWriteF("LS  : L=%2d N=%2d TSOLVE=%10.4g TIT=%10.4g\n", level, lresult->number_of_linear_iterations, ti, ti / lresult->number_of_linear_iterations);
    else
      UserWriteF("LS  : L=%2d N=%2d TSOLVE=%10.4g\n", level, lresult->number_of_linear_iterations, ti);
  }
  if (FreeVD(NP_MG(theNP), hbl, level, np->c)) REP_ERR_RETURN(1);
  return 0;
}

INT GpuLsPostProcess(NP_LINEAR_SOLVER *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPULS *np = (NP_GPULS *)theNP;
  if (np->Iter->PostProcess != NULL)
    if ((*np->Iter->PostProcess)(np->Iter, level, x, b, A, result)) NP_RETURN(1, result[0]);
  np->baselevel = MAX(BOTTOMLEVEL(theNP->base.mg), np->baselevel);
  if (np->m) Release(NP_MG(theNP));
  np->m = NULL;
  return 0;
}

INT GpuLsConstructKind(NP_BASE *theNP, INT kind);
INT GpuLsConstruct(NP_BASE *theNP) { return GpuLsConstructKind(theNP, GPULS_LS); }
INT GpuCgConstruct(NP_BASE *theNP) { return GpuLsConstructKind(theNP, GPULS_CG); }
INT GpuBcgsConstruct(NP_BASE *theNP) { return GpuLsConstructKind(theNP, GPULS_BCGS); }

INT GpuLsConstructKind(NP_BASE *theNP, INT kind)
{
  ((NP_GPULS *)theNP)->kind = kind;
  theNP->Init = GpuLsInit;
  theNP->Display = GpuLsDisplay;
  theNP->Execute = NPLinearSolverExecute;
  NP_LINEAR_SOLVER *np = (NP_LINEAR_SOLVER *)theNP;
  np->PreProcess = GpuLsPreProcess;
  np->Defect = GpuLsDefect;
  np->Residuum = GpuLsResiduum;
  np->Solver = GpuLsSolver;
  np->PostProcess = GpuLsPostProcess;
  return 0;
}


// =========================================================================================================================
// assemble.gpufe  (reference interface: NP_ASSEMBLE np/procs/assemble.h:178-210; the loop it replaces: LocalAssemble
// np/procs/assemble.cc:657-706 and NPLocalAssemblePostMatrix :624 of NP_LOCAL_ASSEMBLE)
// The element kernel, which UG leaves to the application's AssembleLocal, is the device library's built-in one (uggpu_assemble).
// Assemble(level) does what LocalAssemble does -- levels 0..level -- and leaves the same MVALUEs, right-hand side, Dirichlet values
// of x and VECSKIP words in UG's data structures, bit for bit; the device copies stay valid for a solver whose PreProcess runs
// inside this numproc's PreProcess/PostProcess bracket (no second upload of the matrix).
// =========================================================================================================================
gpuls::ElemCoefFn g_fe_coef = NULL;
gpuls::DirichletFn g_fe_dirichlet = NULL;

struct NP_GPUFE {
  NP_ASSEMBLE ass;
  INT problem;          // UGGPU_FE_*
  DOUBLE E, nu;
  DOUBLE source[UGGPU_MAX_BS];
  Mirror *m;
};

INT GpuFeInit(NP_BASE *theNP, INT argc, char **argv)
{
  NP_GPUFE *np = (NP_GPUFE *)theNP;
  char buf[64];
  np->problem = UGGPU_FE_POISSON;
  if (ReadArgvChar("P", buf, argc, argv) == 0) {
    if (strcmp(buf, "elasticity") == 0) np->problem = UGGPU_FE_ELASTICITY;
    else if (strcmp(buf, "poisson") != 0) { PrintErrorMessageF('E', "GpuFeInit", "unknown problem '%s' (poisson | elasticity)", buf); return NP_NOT_ACTIVE; }
  }
  if (ReadArgvDOUBLE("E", &np->E, argc, argv)) np->E = 1.0;
  if (ReadArgvDOUBLE("nu", &np->nu, argc, argv)) np->nu = 0.3;
  for (int i = 0; i < UGGPU_MAX_BS; i++) np->source[i] = 0.0;
  np->source[0] = 1.0;
  for (int i = 0; i < argc; i++)
    if (argv[i][0] == 'f') {
      double f[3] = {0, 0, 0};
      char opt[32];
      int k = sscanf(argv[i], "%31s %lf %lf %lf", opt, &f[0], &f[1], &f[2]);
      if (k >= 2 && strcmp(opt, "f") == 0) for (int j = 0; j < UGGPU_MAX_BS; j++) np->source[j] = j < k - 1 ? f[j] : 0.0;
    }
  return NPAssembleInit(theNP, argc, argv);
}

INT GpuFeDisplay(NP_BASE *theNP)
{
  NP_GPUFE *np = (NP_GPUFE *)theNP;
  NPAssembleDisplay(theNP);
  UserWrite("configuration parameters:\n");
  UserWriteF(DISPLAY_NP_FORMAT_SS, "P", np->problem == UGGPU_FE_ELASTICITY ? "elasticity" : "poisson");
  UserWriteF(DISPLAY_NP_FORMAT_SF, "E", (float)np->E);
  UserWriteF(DISPLAY_NP_FORMAT_SF, "nu", (float)np->nu);
  UserWriteF(DISPLAY_NP_FORMAT_SS, "device", "B200 via libuggpu");
  return 0;
}

INT GpuFePreProcess(NP_ASSEMBLE *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUFE *np = (NP_GPUFE *)theNP;
  np->m = Acquire(NP_MG(theNP));
  if (np->m == NULL) NP_RETURN(1, result[0]);
  return 0;
}

INT GpuFeAssemble(NP_ASSEMBLE *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUFE *np = (NP_GPUFE *)theNP;
  MULTIGRID *mg = NP_MG(theNP);
  Mirror *m = np->m ? Live(np->m, mg) : NULL;
  if (m == NULL) { UserWrite("gpufe: Assemble without PreProcess\n"); NP_RETURN(1, result[0]); }
  std::vector<int64_t> eptr; std::vector<int32_t> erow; std::vector<double> coord, coef, val; std::vector<uint32_t> skip;
  for (int l = 0; l <= level; l++) {
    UserWriteF(" [%d:", l);                                                     // assemble.cc:672
    const int k = Mirror::ix(l);
    gpuls::FlatLevel &f = m->fl[k];
    // the pattern (connections) is the grid manager's; values and VECSKIP are about to be replaced
    if (gpuls::FlattenFlags(mg, l, x, f)) { UserWriteF("gpufe: level %d is not a pure nodal vector format\n", l); NP_RETURN(1, result[0]); }
    if (f.bs > UGGPU_MAX_BS) NP_RETURN(1, result[0]);
    if (gpuls::FlattenMatrix(mg, l, A, f)) { UserWrite("gpufe: cannot flatten the matrix pattern\n"); NP_RETURN(1, result[0]); }
    if (gpuls::FlattenElements(mg, l, eptr, erow, coord, skip, f.bs)) NP_RETURN(1, result[0]);
    m->bs = f.bs; m->xdesc = x;
    m->have_level[k] = 0; m->have_transfer[k] = 0; if (l + 1 < MAXLEVEL) m->have_transfer[k + 1] = 0;
    if (api.uggpu_level_create(m->ctx, m->dl(l), f.n, f.bs)) NP_RETURN(dev_fail("uggpu_level_create"), result[0]);
    if (api.uggpu_level_set_flags(m->ctx, m->dl(l), f.vclass.data(), f.vnclass.data(), f.ctl.data(), NULL)) NP_RETURN(dev_fail("uggpu_level_set_flags"), result[0]);
    if (api.uggpu_mat_set_pattern(m->ctx, m->dl(l), m->handle(A), f.rowptr.data(), f.col.data())) NP_RETURN(dev_fail("uggpu_mat_set_pattern"), result[0]);
    // the application's part of AssembleLocal: coefficients, Dirichlet values of x on the boundary vertices
    const size_t nelem = eptr.size() - 1;
    coef.clear();
    if (g_fe_coef) { coef.reserve(nelem); for (ELEMENT *e = FIRSTELEMENT(GRID_ON_LEVEL(mg, l)); e != NULL; e = SUCCE(e)) coef.push_back((*g_fe_coef)(e)); }
    m->buf.resize((size_t)f.n * f.bs + 1);
    gpuls::GatherVector(mg, l, x, f.bs, m->buf.data());
    for (int r = 0; r < f.n; r++)
      for (int a = 0; a < f.bs; a++)
        if (skip[r] & (1u << a)) m->buf[(size_t)r * f.bs + a] = g_fe_dirichlet ? (*g_fe_dirichlet)(&coord[(size_t)r * DIM], a) : 0.0;
    if (api.uggpu_vec_upload(m->ctx, m->dl(l), m->handle(x), m->buf.data())) NP_RETURN(dev_fail("uggpu_vec_upload"), result[0]);
    gpuls::ScatterVector(mg, l, x, f.bs, m->buf.data());                        // `*sptr[i] = sol[i]`, assemble.cc:692
    if (api.uggpu_vec_alloc(m->ctx, m->dl(l), m->handle(b))) NP_RETURN(dev_fail("uggpu_vec_alloc"), result[0]);
    uggpu_fe_cfg cfg;
    cfg.problem = (int)np->problem; cfg.dim = DIM; cfg.E = np->E; cfg.nu = np->nu;
    for (int a = 0; a < UGGPU_MAX_BS; a++) cfg.source[a] = np->source[a];
    if (api.uggpu_assemble(m->ctx, m->dl(l), m->handle(x), m->handle(b), m->handle(A), &cfg, (int64_t)nelem, eptr.data(), erow.data(),
                           coef.empty() ? NULL : coef.data(), coord.data(), skip.data())) NP_RETURN(dev_fail("uggpu_assemble"), result[0]);
    // results into UG's data structures: MVALUEs, right-hand side, VECSKIP
    val.assign(f.val.size(), 0.0);
    if (api.uggpu_mat_get(m->ctx, m->dl(l), m->handle(A), NULL, NULL, val.data())) NP_RETURN(dev_fail("uggpu_mat_get"), result[0]);
    if (gpuls::ScatterMatrixValues(mg, l, A, val, f.bs) || gpuls::ScatterSkip(mg, l, skip)) NP_RETURN(1, result[0]);
    f.val = val; f.skip = skip;
    if (Download(m, l, b)) NP_RETURN(1, result[0]);
    if (api.uggpu_set_fullrefinelevel(m->ctx, m->dl(FULLREFINELEVEL(mg)))) NP_RETURN(dev_fail("uggpu_set_fullrefinelevel"), result[0]);
    m->have_level[k] = 1; m->level_A[k] = A; m->level_x[k] = x;                 // a solver inside the bracket finds the matrix on the device
    UserWrite("a]");
  }
  UserWrite(" [d]\n");
  return 0;
}

INT GpuFePostProcess(NP_ASSEMBLE *theNP, INT level, VECDATA_DESC *x, VECDATA_DESC *b, MATDATA_DESC *A, INT *result)
{
  NP_GPUFE *np = (NP_GPUFE *)theNP;
  if (np->m != NULL) { if (Live(np->m, NP_MG(theNP))) Release(NP_MG(theNP)); np->m = NULL; }
  return 0;
}

INT GpuFeConstruct(NP_BASE *theNP)
{
  theNP->Init = GpuFeInit;
  theNP->Display = GpuFeDisplay;
  theNP->Execute = NPAssembleExecute;
  NP_ASSEMBLE *np = (NP_ASSEMBLE *)theNP;
  np->PreProcess = GpuFePreProcess;
  np->Assemble = GpuFeAssemble;
  np->PostProcess = GpuFePostProcess;
  return 0;
}

}  // namespace

// ---- savedata / loaddata on the device mirror (gpuls_np.h) ---------------------------------------------------------------------------
namespace {
// node ID -> (level, row) after RenumberMultiGrid, the order of the file body (data_io.cc:906-925); rows = list positions, as uploaded
int NodeMap(MULTIGRID *mg, std::vector<int32_t> &idl, std::vector<int32_t> &idr)
{
  if (RenumberMultiGrid(mg, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0) != GM_OK) { UserWrite("ERROR: cannot renumber multigrid\n"); return 1; }   // data_io.cc:675
  int nn = 0;
  for (int l = 0; l <= TOPLEVEL(mg); l++) nn += NN(GRID_ON_LEVEL(mg, l));
  idl.assign(nn, -1); idr.assign(nn, -1);
  for (int l = 0; l <= TOPLEVEL(mg); l++) {
    GRID *g = GRID_ON_LEVEL(mg, l);
    int r = 0;
    for (VECTOR *v = FIRSTVECTOR(g); v != NULL; v = SUCCVC(v)) VINDEX(v) = r++;
    for (NODE *n = PFIRSTNODE(g); n != NULL; n = SUCCN(n)) {
      const int id = ID(n);
      if (id < 0 || id >= nn || idl[id] >= 0) { UserWrite("internal ERROR: id is out of range\n"); return 1; }                                  // data_io.cc:546
      idl[id] = l; idr[id] = (int32_t)VINDEX(NVECTOR(n));
    }
  }
  return 0;
}
std::string DataFileName(const char *name, const char *type, int number)
{
  std::string f = name;
  if (number != -1) { char b[16]; snprintf(b, sizeof b, ".%06d", number); f += b; }                                                              // data_io.cc:722
  return f + ".ug.data." + type;
}
}

int gpuls::SaveData(MULTIGRID *mg, const char *name, const char *type, int number, double time, double dt, double ndt, int n, VECDATA_DESC **vds)
{
  Mirror *m = Find(mg);
  if (m == NULL || n < 1 || n > 100) { UserWrite("gpuls::SaveData: no device mirror (call inside a PreProcess/PostProcess bracket)\n"); return 1; }
  std::vector<int32_t> idl, idr;
  if (NodeMap(mg, idl, idr)) return 1;
  for (auto &l : idl) if (l >= 0) l += m->loff;          // UG's level numbers -> the device library's
  std::vector<int> vec(n);
  std::vector<std::string> names(n), comps(n);
  std::vector<const char *> np(n), cp(n);
  for (int i = 0; i < n; i++) {
    if (vds[i] == NULL) { UserWrite("gpuls::SaveData: eval procs are not offered\n"); return 1; }
    const int nc = gpuls::NodeComps(vds[i]);
    if (nc <= 0) { PrintErrorMessageF('E', "SaveData", "vd mismatch for io (no %d)", i); return 1; }                                            // data_io.cc:707
    vec[i] = m->handle(vds[i]);
    names[i] = ENVITEM_NAME(vds[i]);
    for (int j = 0; j < nc; j++) comps[i] += vds[i]->compNames[j];
    np[i] = names[i].c_str(); cp[i] = comps[i].c_str();
  }
  uggpu_data_general g;
  char *ident = GetStringVar(":IDENTIFICATION");
  g.ident = ident ? ident : "---";
  g.mgfile = "saved_without_mg";
  g.time = number != -1 ? time : -1.0; g.dt = number != -1 ? dt : -1.0; g.ndt = number != -1 ? ndt : -1.0;
  g.nparfiles = 1; g.me = 0; g.magic_cookie = MG_MAGIC_COOKIE(mg);
  const std::string file = DataFileName(name, type, number);
  if (api.uggpu_savedata(m->ctx, file.c_str(), type, &g, n, vec.data(), np.data(), cp.data(), (int64_t)idl.size(), idl.data(), idr.data())) return dev_fail("uggpu_savedata");
  return 0;
}

int gpuls::LoadData(MULTIGRID *mg, const char *name, const char *type, int number, int n, VECDATA_DESC **vds)
{
  Mirror *m = Find(mg);
  if (m == NULL || n < 1 || n > 100) { UserWrite("gpuls::LoadData: no device mirror (call inside a PreProcess/PostProcess bracket)\n"); return 1; }
  std::vector<int32_t> idl, idr;
  if (NodeMap(mg, idl, idr)) return 1;
  for (auto &l : idl) if (l >= 0) l += m->loff;          // UG's level numbers -> the device library's
  std::vector<int> vec(n);
  for (int i = 0; i < n; i++) vec[i] = vds[i] ? m->handle(vds[i]) : -1;
  uggpu_data_general g;
  const std::string file = DataFileName(name, type, number);
  if (api.uggpu_loaddata(m->ctx, file.c_str(), n, vec.data(), (int64_t)idl.size(), idl.data(), idr.data(), &g)) return dev_fail("uggpu_loaddata");
  if (SetStringValue(":IO:TIME", g.time) || SetStringValue(":IO:DT", g.dt) || SetStringValue(":IO:NDT", g.ndt)) return 1;                       // data_io.cc:500-502
  return 0;
}

void gpuls::SetFEData(gpuls::ElemCoefFn coef, gpuls::DirichletFn dirichlet) { g_fe_coef = coef; g_fe_dirichlet = dirichlet; }

INT NS_DIM_PREFIX InitGpuLS(void)
{
  if (CreateClass(ITER_CLASS_NAME ".gpujac", sizeof(NP_GPUJAC), GpuJacConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ITER_CLASS_NAME ".gpugs", sizeof(NP_GPUJAC), GpuGsConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ITER_CLASS_NAME ".gpusgs", sizeof(NP_GPUJAC), GpuSgsConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ITER_CLASS_NAME ".gpusor", sizeof(NP_GPUJAC), GpuSorConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ITER_CLASS_NAME ".gpuilu", sizeof(NP_GPUJAC), GpuIluConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(TRANSFER_CLASS_NAME ".gputransfer", sizeof(NP_GPUTRANSFER), GpuTransferConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ITER_CLASS_NAME ".gpulmgc", sizeof(NP_GPULMGC), GpuLmgcConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(LINEAR_SOLVER_CLASS_NAME ".gpuls", sizeof(NP_GPULS), GpuLsConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(LINEAR_SOLVER_CLASS_NAME ".gpucg", sizeof(NP_GPULS), GpuCgConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(LINEAR_SOLVER_CLASS_NAME ".gpubcgs", sizeof(NP_GPULS), GpuBcgsConstruct)) REP_ERR_RETURN(__LINE__);
  if (CreateClass(ASSEMBLE_CLASS_NAME ".gpufe", sizeof(NP_GPUFE), GpuFeConstruct)) REP_ERR_RETURN(__LINE__);
  return 0;
}
