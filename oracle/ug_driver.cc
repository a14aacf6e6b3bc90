// oracle/ug_driver.cc -- TEST INFRASTRUCTURE ONLY (golden-vector generator, reference runner and
// CPU-baseline timer).  Never linked into the product; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may execute the binary built from it.
//
// Links against the UNMODIFIED reference compiled from /root/reference (oracle/Makefile) and drives
// it exactly as SURVEY.md Appendix A/B documents: builds a hierarchy through UG's own commands,
// assembles P1/Q1 Poisson or 3x3 linear elasticity into MVALUEs, then
//   --dump F   writes the flattened hierarchy (via the product's PreProcess flattening code,
//              ug_b200/host/gpuls_flatten.cc) plus, with --ops, the result of every individual
//              reference call on the hot path (dmatmul*, BLAS-1, l_jac, Smoother, StandardRestrict,
//              StandardInterpolateCorrection, Lmgc) and, with --solve, the defect history and
//              iterates of `ls`+`lmgc`+`jac`+`transfer`.  These dumps ARE the golden vectors
//              (the reference ships none).
//   --time     times the reference V-cycle and kernels on this host (CPU baseline, 1 core).
//   --gpu      additionally runs the same script with the gpuls numproc family (drop-in test) and
//              compares against the CPU classes in-process.
#include "config.h"
#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <unistd.h>
#include <string>
#include <vector>
#include <map>

#include "initug.h"
#include "gm.h"
#include "np.h"
#include "std_domain.h"
#include "cmdline.h"
#include "ugm.h"
#include "algebra.h"
#include "shapes.h"
#include "evm.h"
#include "numproc.h"
#include "iter.h"
#include "ls.h"
#include "transfer.h"
#include "transgrid.h"
#include "ugdevices.h"
#include "refine.h"
#include "assemble.h"
#include "disctools.h"
#include "data_io.h"

#include "gpuls_flatten.h"
#ifdef WITH_GPULS
#include "gpuls_np.h"
#endif

USING_UG_NAMESPACES
using namespace PPIF;

// ------------------------------------------------------------------------------------------------
// dump file: sequence of records  [u32 namelen][name][u8 dtype][u64 count][raw]   (magic "UGH1\n")
// dtype: 0=i32 1=f64 2=u8 3=u32
struct Dump {
  FILE *f = NULL;
  void open(const char *path) { f = fopen(path, "wb"); if (!f) { perror(path); exit(2); } fwrite("UGH1\n", 1, 5, f); }
  void rec(const std::string &name, int dtype, const void *p, size_t count, size_t esz) {
    if (!f) return;
    uint32_t nl = (uint32_t)name.size(); uint8_t dt = (uint8_t)dtype; uint64_t c = count;
    fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f); fwrite(&dt, 1, 1, f); fwrite(&c, 8, 1, f);
    if (count) fwrite(p, esz, count, f);
  }
  void i32(const std::string &n, const std::vector<int32_t> &v) { rec(n, 0, v.data(), v.size(), 4); }
  void f64(const std::string &n, const std::vector<double> &v) { rec(n, 1, v.data(), v.size(), 8); }
  void u8(const std::string &n, const std::vector<uint8_t> &v) { rec(n, 2, v.data(), v.size(), 1); }
  void u32(const std::string &n, const std::vector<uint32_t> &v) { rec(n, 3, v.data(), v.size(), 4); }
  void scalar_i(const std::string &n, int v) { int32_t x = v; rec(n, 0, &x, 1, 4); }
  void scalar_d(const std::string &n, double v) { rec(n, 1, &v, 1, 8); }
  void close() { if (f) fclose(f); f = NULL; }
};
static Dump D;

#include <execinfo.h>
#include <signal.h>
// a crash inside the reference: print where (addr2line -e ugoracleN <addresses>) instead of dying silently
static void on_crash(int sig)
{
  void *bt[48];
  int n = backtrace(bt, 48);
  fprintf(stderr, "ug_driver: signal %d inside the reference; backtrace:\n", sig);
  backtrace_symbols_fd(bt, n, 2);
  fflush(NULL);
  _exit(128 + sig);
}

static void cmd(const char *fmt, ...)
{
  char b[2048];
  va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
  if (ExecCommand(b)) { fprintf(stderr, "UG command failed: %s\n", b); exit(3); }
}

static double now()
{
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static INT BndCond(void *, void *, DOUBLE *, DOUBLE *v, INT *t) { v[0] = 0; *t = 1; return 0; }

// ------------------------------------------------------------------------------------------------
struct Opt {
  std::string grid = (DIM == 3) ? "tet" : "tri";
  int bs = 1, refine = 3, adapt = 0, nu1 = 2, nu2 = 2, gamma = 1, cycles = 10, reps = 5;
  double damp = (DIM == 3) ? 0.6 : 0.8;
  std::string dump, gpu, barrier_dir;   // barrier_dir: --replicas rendezvous (see time_reference)
  std::string smoother = "jac";         // reference smoother class used by lmgc: jac | gs | sgs | sor | ilu (iter.cc:10343-10366)
  double beta = 0.0;                    // ilu $beta (iter.cc:5415): diagonal modification of l_ilubthdecomp
  int baselevel = 0;                    // lmgc $b
  int barrier_n = 0, barrier_id = 0;
  bool ops = false, solve = false, timeit = false, quiet = true, nokrylov = false, elems = false;
  bool imat = false;                    // transfer $M: RestrictByMatrix / InterpolateCorrectionByMatrix on stored interpolation matrices
  bool galerkin = false;                // --galerkin (with --imat): Galerkin coarse-grid operators by AssembleGalerkinByMatrix, cascaded from the top level down
  std::string savedata;
  bool assemble = false;                // --assemble: run the reference's LocalAssemble (np/procs/assemble.cc:657) with the element kernel below and dump what it leaves
  bool lean = false;                    // --lean: dumps without the BLAS-1/2 and transfer records, the coordinates and the Krylov runs
  // --amg CLASS "INIT": the cycle continues below level 0 on algebraic levels built by the reference's own AMG transfer numproc
  // (np/procs/amgtransfer.cc: classes selectionAMG / clusterAMG), attached to the transfer class with `$amg` (transfer.cc:593, :660)
  std::string amg_class, amg_init;
  // --gpuamg "OPTIONS": the gputransfer numprocs of the --gpu run get `$gpuamg OPTIONS` (the algebraic levels are built by the device
  // library itself and exist on the device only) instead of `$amg amgt`; the CPU side keeps the reference's AMG numproc (--amg)
  std::string gpuamg;
  bool hooks = false;                   // --hooks (with --gpu): InterpolateNewVectors / ProjectSolution of transfer vs gputransfer on the same vectors
  bool transferD = false;               // --transferD: transfer $D (AssembleDirichletBoundary on every level in the transfer's PreProcess, transfer.cc:666)
  bool levelopt = false;                // --levelopt: transfer $L, level optimisation after every level's post-smoothing (transfer.cc:574, :812, MinimizeLevel :488)
  bool collapse = false;                // --collapse: after the --refine steps the surface becomes level 0 (UG's `collapse`, gm/ugm.cc:3930): a large level 0 for the AMG
  int refine2 = 0;                      // --refine2 K: K uniform refinements after the collapse
};

static MULTIGRID *mg;
static VECDATA_DESC *vx, *vb, *vc, *vt;
static MATDATA_DESC *mA;
static int BS;

// LO: the bottom level of the multigrid while a dump is written -- 0, or negative once the algebraic levels of an AMG transfer exist
// (--amg with $hold).  Dumps number their levels from 0: record "L<k>/..." belongs to UG's level LO + k.
static int LO = 0;
static std::string L(const char *what, int lev) { char b[128]; snprintf(b, sizeof b, "L%d/%s", lev - LO, what); return b; }

// deterministic test vectors: k * 2^-20 from a 64-bit LCG, seed mixes level+tag
static void fill_lcg(const VECDATA_DESC *vd, int lev, uint64_t seed)
{
  uint64_t s = 12345ull + seed * 0x9E3779B97F4A7C15ull + (uint64_t)lev * 1000003ull;
  for (VECTOR *v = FIRSTVECTOR(GRID_ON_LEVEL(mg, lev)); v != NULL; v = SUCCVC(v))
    for (int i = 0; i < BS; i++) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      int64_t k = (int64_t)((s >> 33) & 0xFFFFF) - 0x80000;
      VVALUE(v, VD_CMP_OF_TYPE(vd, VTYPE(v), i)) = (double)k * (1.0 / 1048576.0);
    }
}

static std::vector<double> gather(const VECDATA_DESC *vd, int lev)
{
  std::vector<double> h((size_t)NVEC(GRID_ON_LEVEL(mg, lev)) * BS);
  gpuls::GatherVector(mg, lev, vd, BS, h.data());
  return h;
}
static void dumpvec(const std::string &name, const VECDATA_DESC *vd, int lev) { D.f64(L(name.c_str(), lev), gather(vd, lev)); }

// ------------------------------------------------------------------------------------------------
// element matrices
static const double HEXLOC[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const double QUADLOC[4][2] = {{0,0},{1,0},{1,1},{0,1}};

// shape function gradients in reference coordinates for tensor-product elements
static void tp_shape(int nc, const double *xi, double *N, double (*dN)[DIM])
{
  for (int i = 0; i < nc; i++) {
    double f[DIM], df[DIM];
    for (int d = 0; d < DIM; d++) {
      double a = (DIM == 3) ? HEXLOC[i][d] : QUADLOC[i][d % 2];
      f[d] = a ? xi[d] : 1.0 - xi[d];
      df[d] = a ? 1.0 : -1.0;
    }
    N[i] = 1.0;
    for (int d = 0; d < DIM; d++) N[i] *= f[d];
    for (int d = 0; d < DIM; d++) {
      dN[i][d] = df[d];
      for (int e = 0; e < DIM; e++) if (e != d) dN[i][d] *= f[e];
    }
  }
}

static double det_inv(const double J[DIM][DIM], double Ji[DIM][DIM])
{
#if DIM == 2
  double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
  return det;
#else
  double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  return det;
#endif
}

// Quadrature points: simplex -> centroid rule is exact for P1 stiffness; tensor elements -> 2-point Gauss.
// Returns gradients G[q][i][d], weights*detJ W[q], shape values N[q][i].
struct QP { double w; double N[8]; double G[8][DIM]; };
static int element_qps(ELEMENT *e, QP *qp)
{
  int nc = CORNERS_OF_ELEM(e);
  double X[8][DIM];
  for (int i = 0; i < nc; i++) for (int d = 0; d < DIM; d++) X[i][d] = CVECT(MYVERTEX(CORNER(e, i)))[d];
  if (nc == DIM + 1) {
    // barycentric: N_0 = 1 - sum xi, N_i = xi_{i-1}
    double J[DIM][DIM], Ji[DIM][DIM];
    for (int d = 0; d < DIM; d++) for (int k = 0; k < DIM; k++) J[k][d] = X[k + 1][d] - X[0][d];  // J[k][d] = dx_d/dxi_k
    double det = det_inv(J, Ji);
    double vol = fabs(det) / ((DIM == 3) ? 6.0 : 2.0);
    qp[0].w = vol;
    for (int i = 0; i < nc; i++) qp[0].N[i] = 1.0 / nc;
    for (int d = 0; d < DIM; d++) {
      double s = 0;
      for (int k = 0; k < DIM; k++) { qp[0].G[k + 1][d] = Ji[d][k]; s += Ji[d][k]; }  // grad_x xi_k
      qp[0].G[0][d] = -s;
    }
    return 1;
  }
  const double g[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
  int nq = 0;
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < ((DIM == 3) ? 2 : 1); c++) {
    double xi[3] = {g[a], g[b], g[c]};
    double dN[8][DIM], J[DIM][DIM], Ji[DIM][DIM];
    tp_shape(nc, xi, qp[nq].N, dN);
    for (int k = 0; k < DIM; k++) for (int d = 0; d < DIM; d++) { J[k][d] = 0; for (int i = 0; i < nc; i++) J[k][d] += dN[i][k] * X[i][d]; }
    double det = det_inv(J, Ji);
    qp[nq].w = fabs(det) / ((DIM == 3) ? 8.0 : 4.0);
    for (int i = 0; i < nc; i++) for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < DIM; k++) s += Ji[d][k] * dN[i][k]; qp[nq].G[i][d] = s; }
    nq++;
  }
  return nq;
}

static void assemble(void)
{
  int t0 = 0;
  const double E = 1.0, nu = 0.3, lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
  for (int l = 0; l <= TOPLEVEL(mg); l++) {
    GRID *g = GRID_ON_LEVEL(mg, l);
    for (VECTOR *v = FIRSTVECTOR(g); v; v = SUCCVC(v)) {
      t0 = VTYPE(v);
      VECSKIP(v) = 0;
      for (int i = 0; i < BS; i++) {
        VVALUE(v, VD_CMP_OF_TYPE(vx, t0, i)) = 0; VVALUE(v, VD_CMP_OF_TYPE(vb, t0, i)) = 0;
        VVALUE(v, VD_CMP_OF_TYPE(vc, t0, i)) = 0; VVALUE(v, VD_CMP_OF_TYPE(vt, t0, i)) = 0;
      }
      for (MATRIX *m = VSTART(v); m; m = MNEXT(m))
        for (int k = 0; k < BS * BS; k++) MVALUE(m, MD_MCMP_OF_RT_CT(mA, t0, t0, k)) = 0;
    }
    QP qp[8];
    for (ELEMENT *e = FIRSTELEMENT(g); e; e = SUCCE(e)) {
      int nc = CORNERS_OF_ELEM(e);
      int nq = element_qps(e, qp);
      for (int i = 0; i < nc; i++) {
        VECTOR *vi = NVECTOR(CORNER(e, i));
        for (int q = 0; q < nq; q++) {
          if (BS == 1) VVALUE(vi, VD_CMP_OF_TYPE(vb, t0, 0)) += qp[q].w * qp[q].N[i];
          else VVALUE(vi, VD_CMP_OF_TYPE(vb, t0, BS - 1)) += -qp[q].w * qp[q].N[i];  // gravity on last comp
        }
        for (int j = 0; j < nc; j++) {
          VECTOR *vj = NVECTOR(CORNER(e, j));
          MATRIX *m = GetMatrix(vi, vj);
          if (!m) { fprintf(stderr, "missing connection\n"); exit(4); }
          for (int q = 0; q < nq; q++) {
            const double *gi = qp[q].G[i], *gj = qp[q].G[j];
            double dot = 0; for (int d = 0; d < DIM; d++) dot += gi[d] * gj[d];
            if (BS == 1) MVALUE(m, MD_MCMP_OF_RT_CT(mA, t0, t0, 0)) += qp[q].w * dot;
            else
              for (int a = 0; a < BS; a++) for (int b = 0; b < BS; b++) {
                double k = lam * gi[a] * gj[b] + mu * gi[b] * gj[a] + ((a == b) ? mu * dot : 0.0);
                MVALUE(m, MD_MCMP_OF_RT_CT(mA, t0, t0, a * BS + b)) += qp[q].w * k;
              }
          }
        }
      }
    }
    // Dirichlet: identity rows + skip flags on boundary vertices (Appendix B)
    for (NODE *n = FIRSTNODE(g); n; n = SUCCN(n))
      if (OBJT(MYVERTEX(n)) == BVOBJ) {
        VECTOR *v = NVECTOR(n);
        VECSKIP(v) = (1u << BS) - 1;
        for (int i = 0; i < BS; i++) VVALUE(v, VD_CMP_OF_TYPE(vb, t0, i)) = 0;
        for (MATRIX *m = VSTART(v); m; m = MNEXT(m))
          for (int k = 0; k < BS * BS; k++) MVALUE(m, MD_MCMP_OF_RT_CT(mA, t0, t0, k)) = 0;
        for (int i = 0; i < BS; i++) MVALUE(VSTART(v), MD_MCMP_OF_RT_CT(mA, t0, t0, i * BS + i)) = 1.0;
      }
  }
}


// ------------------------------------------------------------------------------------------------
// Element-loop assembly through the reference's own NP_LOCAL_ASSEMBLE machinery: np/procs/assemble.cc:657-706
// LocalAssemble (dset / dmatset / CLEAR_VECSKIP, elements in list order, GetElementVVMPtrs np/udm/disctools.cc:1113, `+=` of the local
// defect and matrix, SetElementDirichletFlags :1763 on boundary elements) and :624 NPLocalAssemblePostMatrix (AssembleDirichletBoundary
// disctools.cc:1837 on every level).  UG leaves the element kernel to the application (AssembleLocal); class `fe` below is that
// application: P1 / Q1 diffusion with one coefficient per element, or isotropic linear elasticity, source term, Dirichlet values g(x)
// on the whole boundary.  The local matrix is summed over the quadrature points first and added to the global one once per element
// (assemble() above adds every quadrature term directly; for simplices -- one point -- the two agree to the bit when the coefficient is 1).
// assemble.cc:606-731 defines these four WITHOUT the namespace prefix its header declares them with: the library exports them in the
// global namespace, so they are declared (and called) there
INT NPLocalAssembleInit(NP_LOCAL_ASSEMBLE *, INT, char **);
INT NPLocalAssembleDisplay(NP_LOCAL_ASSEMBLE *);
INT NPLocalAssembleConstruct(NP_ASSEMBLE *);
INT NPLocalAssemblePostMatrix(NP_LOCAL_ASSEMBLE *, INT, VECDATA_DESC *, VECDATA_DESC *, MATDATA_DESC *, INT *);
static void restore_problem(void);
static DOUBLE fe_sol[24], fe_def[24], fe_mat[24 * 24];
static INT fe_vecskip[24];
static const double fe_E = 1.0, fe_nu = 0.3;

static double fe_coef(ELEMENT *e)
{
  int nc = CORNERS_OF_ELEM(e);
  double c[3] = {0, 0, 0};
  for (int i = 0; i < nc; i++) for (int d = 0; d < DIM; d++) c[d] += CVECT(MYVERTEX(CORNER(e, i)))[d];
  for (int d = 0; d < DIM; d++) c[d] /= nc;
  return 1.0 + 0.5 * c[0] + 0.25 * c[1] * c[1] + (DIM == 3 ? 0.125 * c[2] : 0.0);
}
static void fe_source(double *f) { for (int a = 0; a < BS; a++) f[a] = 0.0; if (BS == 1) f[0] = 1.0; else f[BS - 1] = -1.0; }
static double fe_dirichlet(const double *x, int a) { return (a + 1) * (0.25 * x[0] - 0.5 * x[1] + (DIM == 3 ? 0.125 * x[2] : 0.0)); }

static INT FEPreProcess(NP_LOCAL_ASSEMBLE *, INT, VECDATA_DESC *, VECDATA_DESC *, MATDATA_DESC *, DOUBLE **sol, DOUBLE **def, DOUBLE **mat, INT **vecskip, INT *)
{
  *sol = fe_sol; *def = fe_def; *mat = fe_mat; *vecskip = fe_vecskip;
  return 0;
}

static INT FEAssembleLocal(ELEMENT *e, INT *result)
{
  const int nc = CORNERS_OF_ELEM(e), m = nc * BS;
  if (nc != DIM + 1 && nc != (1 << DIM)) { result[0] = __LINE__; return 1; }
  const double lam = fe_E * fe_nu / ((1 + fe_nu) * (1 - 2 * fe_nu)), mu = fe_E / (2 * (1 + fe_nu));
  QP qp[8];
  const int nq = element_qps(e, qp);
  const double kappa = fe_coef(e);
  double f[3]; fe_source(f);
  for (int q = 0; q < nq; q++) {
    const double wk = kappa * qp[q].w;
    for (int i = 0; i < nc; i++) {
      const double wn = qp[q].w * qp[q].N[i];
      for (int a = 0; a < BS; a++) fe_def[i * BS + a] += wn * f[a];
    }
    for (int i = 0; i < nc; i++)
      for (int j = 0; j < nc; j++) {
        const double *gi = qp[q].G[i], *gj = qp[q].G[j];
        double dot = 0; for (int d = 0; d < DIM; d++) dot += gi[d] * gj[d];
        if (BS == 1) fe_mat[i * m + j] += wk * dot;
        else
          for (int a = 0; a < BS; a++) for (int b = 0; b < BS; b++) {
            double k = lam * gi[a] * gj[b] + mu * gi[b] * gj[a] + ((a == b) ? mu * dot : 0.0);
            fe_mat[(i * BS + a) * m + j * BS + b] += wk * k;
          }
      }
  }
  if (OBJT(e) == BEOBJ)
    for (int i = 0; i < nc; i++)
      if (OBJT(MYVERTEX(CORNER(e, i))) == BVOBJ)
        for (int a = 0; a < BS; a++) { fe_vecskip[i * BS + a] = 1; fe_sol[i * BS + a] = fe_dirichlet(CVECT(MYVERTEX(CORNER(e, i))), a); }
  return 0;
}

static INT FEInit(NP_BASE *theNP, INT argc, char **argv) { return ::NPLocalAssembleInit((NP_LOCAL_ASSEMBLE *)theNP, argc, argv); }
static INT FEDisplay(NP_BASE *theNP) { return ::NPLocalAssembleDisplay((NP_LOCAL_ASSEMBLE *)theNP); }
static INT FEConstruct(NP_BASE *theNP)
{
  theNP->Init = FEInit; theNP->Display = FEDisplay; theNP->Execute = NPAssembleExecute;
  NP_LOCAL_ASSEMBLE *la = (NP_LOCAL_ASSEMBLE *)theNP;
  ::NPLocalAssembleConstruct(&la->assemble);
  la->PreProcess = FEPreProcess; la->AssembleLocal = FEAssembleLocal; la->AssembleLocalDefect = NULL; la->AssembleLocalMatrix = NULL;
  la->PostMatrix = ::NPLocalAssemblePostMatrix; la->PostProcess = NULL;
  return 0;
}

// what the reference's LocalAssemble leaves on every level: matrix values (canonical entry order), right-hand side, solution (random
// values, g(x) on the boundary), VECSKIP, and the per-element coefficients the element kernel used.  The elements and coordinates are
// in the hierarchy part of the dump (--assemble implies --elems).  Afterwards the problem of the other records is restored.
static NP_ASSEMBLE *run_fe_assemble(const char *cls, bool keep_open = false)
{
  static int made = 0;
  int top = TOPLEVEL(mg);
  INT result = 0;
  if (!made && CreateClass(ASSEMBLE_CLASS_NAME ".fe", sizeof(NP_LOCAL_ASSEMBLE), FEConstruct)) { fprintf(stderr, "CreateClass fe failed\n"); exit(12); }
  made = 1;
  char nm[32]; snprintf(nm, sizeof nm, "ass_%s", cls);
  NP_ASSEMBLE *ass = (NP_ASSEMBLE *)GetNumProcByName(mg, nm, ASSEMBLE_CLASS_NAME);
  if (!ass) {
    cmd("npcreate %s $c %s", nm, cls);
    if (strcmp(cls, "fe") == 0) cmd("npinit %s $A MAT $x sol $b rhs", nm);
    else if (BS == 1) cmd("npinit %s $A MAT $x sol $b rhs $P poisson $f 1", nm);
    else cmd("npinit %s $A MAT $x sol $b rhs $P elasticity $E %.17g $nu %.17g $f %s-1", nm, fe_E, fe_nu, BS == 3 ? "0 0 " : "0 ");
    ass = (NP_ASSEMBLE *)GetNumProcByName(mg, nm, ASSEMBLE_CLASS_NAME);
  }
  if (!ass) { fprintf(stderr, "numproc %s not found\n", nm); exit(12); }
  for (int l = 0; l <= top; l++) fill_lcg(vx, l, 7);
  if ((*ass->PreProcess)(ass, top, vx, vb, mA, &result)) { fprintf(stderr, "%s: PreProcess failed (%d)\n", cls, (int)result); exit(12); }
  if ((*ass->Assemble)(ass, top, vx, vb, mA, &result)) { fprintf(stderr, "%s: Assemble failed (%d)\n", cls, (int)result); exit(12); }
  if (!keep_open && (*ass->PostProcess)(ass, top, vx, vb, mA, &result)) { fprintf(stderr, "%s: PostProcess failed (%d)\n", cls, (int)result); exit(12); }
  return ass;
}

static void dump_assemble(const Opt &o)
{
  (void)o;
  int top = TOPLEVEL(mg);
  run_fe_assemble("fe");
  D.scalar_d("asm/E", fe_E); D.scalar_d("asm/nu", fe_nu);
  { double f[3]; fe_source(f); D.rec("asm/source", 1, f, BS, 8); }
  for (int l = 0; l <= top; l++) {
    gpuls::FlatLevel f;
    if (gpuls::FlattenFlags(mg, l, vx, f) || gpuls::FlattenMatrix(mg, l, mA, f)) { fprintf(stderr, "FlattenMatrix(assemble) failed\n"); exit(12); }
    D.f64(L("asm/val", l), f.val); D.u32(L("asm/skip", l), f.skip);
    dumpvec("asm/rhs", vb, l); dumpvec("asm/sol", vx, l);
    std::vector<double> coef;
    for (ELEMENT *e = FIRSTELEMENT(GRID_ON_LEVEL(mg, l)); e; e = SUCCE(e)) coef.push_back(fe_coef(e));
    D.f64(L("asm/coef", l), coef);
  }
  restore_problem();
}


// savedata / loaddata: the reference's SaveData (np/udm/data_io.cc:650) writes vectors `sol` and `rhs` (seeded values) of all
// levels without a multigrid file, in binary and ASCII mode; the dump records the node-ID order the file body follows.
static void dump_savedata(const Opt &o)
{
  int top = TOPLEVEL(mg);
  for (int l = 0; l <= top; l++) { fill_lcg(vx, l, 11); fill_lcg(vb, l, 12); }
  VECDATA_DESC *vds[2] = {vx, vb};
  // Built with UG_USE_SYSTEM_HEAP (oracle/Makefile) the multigrid heap is a stub of MIN_HEAP_SIZE bytes (gm/ugm.cc:3165) while every
  // allocation is a malloc; SaveData / LoadData size their buffers from `size - used` of that stub (data_io.cc:886, :570) and give up.
  // The size FIELD is raised for the duration of the calls -- no reference source is touched, the buffers still come from malloc.
  const MEM heap_size_field = MGHEAP(mg)->size;
  MGHEAP(mg)->size = MGHEAP(mg)->used + (MEM)(1u << 30);
  EVALUES *ev[2] = {NULL, NULL}; EVECTOR *evec[2] = {NULL, NULL};
  for (const char *type : {"bin", "asc"}) {
    std::string name = o.savedata;
    if (SaveData(mg, (char *)name.c_str(), 1, 1, (char *)type, -1, 0.0, 0.0, 0.0, 2, vds, ev, evec, NULL)) { fprintf(stderr, "SaveData(%s) failed\n", type); exit(13); }
  }
  MGHEAP(mg)->size = heap_size_field;
  for (const char *type : {"bin", "asc"}) {           // the files themselves travel inside the dump
    std::string path = o.savedata + ".ug.data." + type;
    std::vector<uint8_t> bytes;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { perror(path.c_str()); exit(13); }
    for (int c; (c = fgetc(f)) != EOF;) bytes.push_back((uint8_t)c);
    fclose(f);
    D.u8(std::string("savedata/file_") + type, bytes);
    remove(path.c_str());
  }
  std::vector<int32_t> idl, idr;
  int nn = 0;
  for (int l = 0; l <= top; l++) nn += NN(GRID_ON_LEVEL(mg, l));
  idl.assign(nn, -1); idr.assign(nn, -1);
  for (int l = 0; l <= top; l++) {
    gpuls::FlatLevel f;
    if (gpuls::FlattenFlags(mg, l, vx, f)) exit(13);
    for (NODE *n = PFIRSTNODE(GRID_ON_LEVEL(mg, l)); n; n = SUCCN(n)) { idl[ID(n)] = l; idr[ID(n)] = VINDEX(NVECTOR(n)); }
    dumpvec("savedata/sol", vx, l); dumpvec("savedata/rhs", vb, l);
  }
  D.i32("savedata/id_level", idl); D.i32("savedata/id_row", idr);
  D.scalar_i("savedata/magic_cookie", MG_MAGIC_COOKIE(mg));
  { std::vector<uint8_t> cn; for (int j = 0; j < BS; j++) cn.push_back((uint8_t)vx->compNames[j]); D.u8("savedata/compnames", cn); }
  restore_problem();
}

// ------------------------------------------------------------------------------------------------
static void build_hierarchy(const Opt &o)
{
  CreateBoundaryValueProblem("theBVP", BndCond, 0, NULL, 0, NULL);
#if DIM == 3
  cmd("configure theBVP $d Hexahedron");
#else
  cmd("configure theBVP $d Quadrilateral");
#endif
  cmd("newformat F $V n%d: vt 16 $M implicit(vt): mt 2 $I n%d", o.bs, o.bs * o.bs);
  cmd("new themg $b theBVP $f F $h 30000M");
  mg = GetMultigrid((char *)"themg");
#if DIM == 3
  if (o.grid == "tet") {
    cmd("ie 0 1 2 6"); cmd("ie 0 2 3 6"); cmd("ie 0 3 7 6"); cmd("ie 0 7 4 6"); cmd("ie 0 4 5 6"); cmd("ie 0 5 1 6");
  } else cmd("ie 0 1 2 3 4 5 6 7");
#else
  if (o.grid == "tri") { cmd("ie 0 1 2"); cmd("ie 0 2 3"); } else cmd("ie 0 1 2 3");
#endif
  cmd("fixcoarsegrid");
  for (int i = 0; i < o.refine; i++) cmd("refine $a");
  if (o.collapse) { cmd("collapse"); for (int i = 0; i < o.refine2; i++) cmd("refine $a"); }
  for (int r = 0; r < o.adapt; r++) {
    for (int l = 0; l <= TOPLEVEL(mg); l++)
      for (ELEMENT *e = FIRSTELEMENT(GRID_ON_LEVEL(mg, l)); e; e = SUCCE(e)) {
        if (!EstimateHere(e)) continue;
        int nc = CORNERS_OF_ELEM(e); bool in = true;
        for (int d = 0; d < DIM; d++) { double c = 0; for (int i = 0; i < nc; i++) c += CVECT(MYVERTEX(CORNER(e, i)))[d]; if (c / nc > 0.5) in = false; }
        if (in) MarkForRefinement(e, RED, 0);
      }
    cmd("refine");
  }
  cmd("createvector sol rhs cor tmp");
  cmd("creatematrix MAT");
  vx = GetVecDataDescByName(mg, (char *)"sol"); vb = GetVecDataDescByName(mg, (char *)"rhs");
  vc = GetVecDataDescByName(mg, (char *)"cor"); vt = GetVecDataDescByName(mg, (char *)"tmp");
  mA = GetMatDataDescByName(mg, (char *)"MAT");
  if (!vx || !vb || !vc || !vt || !mA) { fprintf(stderr, "descriptor lookup failed\n"); exit(5); }
  BS = o.bs;
  assemble();
}

static void make_numprocs(const Opt &o, const char *pfx, const char *jac, const char *lmgc, const char *transfer, const char *ls, int maxit)
{
  cmd("npcreate %ssmooth $c %s", pfx, jac);
  if (o.smoother == "ilu") cmd("npinit %ssmooth $damp %.17g $beta %.17g", pfx, o.damp, o.beta);
  else cmd("npinit %ssmooth $damp %.17g", pfx, o.damp);
  cmd("npcreate %sbaseit $c lu", pfx);           cmd("npinit %sbaseit", pfx);
  cmd("npcreate %sbasesolver $c ls", pfx);       cmd("npinit %sbasesolver $red 1e-8 $m 10 $I %sbaseit", pfx, pfx);
  if (!o.amg_class.empty()) { cmd("npcreate %samgt $c %s", pfx, o.amg_class.c_str()); cmd("npinit %samgt %s", pfx, o.amg_init.c_str()); }
  cmd("npcreate %stransfer $c %s", pfx, transfer);
  const std::string topt = std::string(o.imat ? " $M" : "") + (o.levelopt ? " $L" : "") + (o.transferD ? " $D" : "");
  if (!o.gpuamg.empty() && strcmp(transfer, "gputransfer") == 0) cmd("npinit %stransfer%s $gpuamg %s", pfx, topt.c_str(), o.gpuamg.c_str());
  else if (!o.amg_class.empty()) cmd("npinit %stransfer%s $amg %samgt", pfx, topt.c_str(), pfx);
  else cmd("npinit %stransfer%s", pfx, topt.c_str());
  cmd("npcreate %slmgc $c %s", pfx, lmgc);
  cmd("npinit %slmgc $S %ssmooth %ssmooth %sbasesolver $T %stransfer $n1 %d $n2 %d $g %d $b %d", pfx, pfx, pfx, pfx, pfx, o.nu1, o.nu2, o.gamma, o.baselevel);
  cmd("npcreate %smgs $c %s", pfx, ls);
  cmd("npinit %smgs $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I %slmgc $display %s", pfx, maxit, pfx, o.quiet ? "no" : "full");
}

// ------------------------------------------------------------------------------------------------
static void dump_hierarchy(const Opt &o, std::vector<gpuls::FlatLevel> &fl)
{
  int top = TOPLEVEL(mg);
  fl.resize(top - LO + 1);
  D.scalar_i("dim", DIM); D.scalar_i("bs", BS); D.scalar_i("toplevel", top - LO);
  D.scalar_i("fullrefinelevel", FULLREFINELEVEL(mg) - LO);
  D.scalar_i("bottomlevel", LO);
  if (o.levelopt) D.scalar_i("level_opt", 1);
  if (!o.amg_class.empty()) {        // which AMG built the algebraic levels: class and init string, as bytes
    D.u8("amg/class", std::vector<uint8_t>(o.amg_class.begin(), o.amg_class.end()));
    D.u8("amg/init", std::vector<uint8_t>(o.amg_init.begin(), o.amg_init.end()));
  }
  D.scalar_d("damp", o.damp); D.scalar_i("nu1", o.nu1); D.scalar_i("nu2", o.nu2); D.scalar_i("gamma", o.gamma);
  D.scalar_i("baselevel", o.baselevel);
  D.scalar_i("smoother", o.smoother == "jac" ? 0 : o.smoother == "gs" ? 1 : o.smoother == "sgs" ? 2 : o.smoother == "sor" ? 3 : 4);
  if (o.smoother == "ilu") D.scalar_d("ilu_beta", o.beta);
  for (int l = LO; l <= top; l++) {
    if (gpuls::FlattenFlags(mg, l, vx, fl[l - LO])) { fprintf(stderr, "FlattenFlags failed\n"); exit(6); }
    if (gpuls::FlattenMatrix(mg, l, mA, fl[l - LO])) { fprintf(stderr, "FlattenMatrix failed\n"); exit(6); }
  }
  D.scalar_i("transfer_mode", o.imat ? 1 : 0);
  for (int l = LO + 1; l <= top; l++) {
    const bool imat = o.imat || l < 1;        // levels < 1 always use the stored interpolation matrices (transfer.cc:733, :756)
    if (imat ? gpuls::FlattenTransferIMAT(mg, l, fl[l - LO]) : gpuls::FlattenTransfer(mg, l, fl[l - LO])) { fprintf(stderr, "FlattenTransfer failed\n"); exit(6); }
    if (LO < 0) D.scalar_i(L("transfer_mode", l), imat ? 1 : 0);
  }
  for (int l = LO; l <= top; l++) {
    gpuls::FlatLevel &f = fl[l - LO];
    D.scalar_i(L("n", l), f.n);
    D.i32(L("rowptr", l), f.rowptr); D.i32(L("col", l), f.col); D.f64(L("val", l), f.val);
    D.u8(L("vclass", l), f.vclass); D.u8(L("vnclass", l), f.vnclass); D.u8(L("ctl", l), f.ctl); D.u32(L("skip", l), f.skip);
    if (l > LO) {
      D.i32(L("p_rowptr", l), f.p_rowptr); D.i32(L("p_col", l), f.p_col); D.f64(L("p_w", l), f.p_w);
      D.i32(L("r_rowptr", l), f.r_rowptr); D.i32(L("r_col", l), f.r_col); D.f64(L("r_w", l), f.r_w);
      D.i32(L("node_row", l), f.node_row);
    }
    dumpvec("rhs", vb, l);
    if (o.elems) {
      // elements in list order: the rows of their corner vectors, and the position of the father element in the list of the level below
      // (-1 on level 0) -- what parallel/dddif/lbrcb.cc works on (centres of mass of the level-0 elements, sons inherit)
      GRID *g = GRID_ON_LEVEL(mg, l);
      std::vector<int32_t> eptr(1, 0), enodes, efather;
      std::map<ELEMENT *, int> pos;
      if (l > LO) { int k = 0; for (ELEMENT *e = FIRSTELEMENT(GRID_ON_LEVEL(mg, l - 1)); e; e = SUCCE(e)) pos[e] = k++; }
      for (ELEMENT *e = FIRSTELEMENT(g); e; e = SUCCE(e)) {
        for (int i = 0; i < CORNERS_OF_ELEM(e); i++) enodes.push_back((int32_t)VINDEX(NVECTOR(CORNER(e, i))));
        eptr.push_back((int32_t)enodes.size());
        efather.push_back(l > 0 && EFATHER(e) ? pos[EFATHER(e)] : -1);
      }
      D.i32(L("elem_ptr", l), eptr); D.i32(L("elem_nodes", l), enodes); D.i32(L("elem_father", l), efather);
    }
    // vertex coordinates in row order (lets tests relate UG's ordering to the synthetic generator)
    if ((o.lean && !o.elems) || l < 0) continue;      // algebraic levels have no nodes
    std::vector<double> xyz((size_t)f.n * DIM);
    for (NODE *n = FIRSTNODE(GRID_ON_LEVEL(mg, l)); n; n = SUCCN(n))
      for (int d = 0; d < DIM; d++) xyz[(size_t)VINDEX(NVECTOR(n)) * DIM + d] = CVECT(MYVERTEX(n))[d];
    D.f64(L("xyz", l), xyz);
  }
}

// every individual reference call on the hot path, inputs and outputs dumped
static void dump_ops(const Opt &o)
{
  int top = TOPLEVEL(mg);
  VEC_SCALAR damp, one, a3;
  for (int i = 0; i < MAX_VEC_COMP; i++) { damp[i] = o.damp; one[i] = 1.0; a3[i] = 0.25 + 0.5 * i; }
  NP_ITER *smooth = (NP_ITER *)GetNumProcByName(mg, "smooth", ITER_CLASS_NAME);
  INT result = 0, bl = 0;
  for (int l = 0; l <= top; l++) {
    GRID *g = GRID_ON_LEVEL(mg, l);
    fill_lcg(vx, l, 1); fill_lcg(vb, l, 2); fill_lcg(vc, l, 3); fill_lcg(vt, l, 4);
    dumpvec("in/x", vx, l); dumpvec("in/b", vb, l); dumpvec("in/c", vc, l); dumpvec("in/t", vt, l);
    // Gauss-Seidel family: triangular solves in VINDEX order, ugiter.cc:412 / :735 / :1343 / :1563
    if (l_setindex(g)) { fprintf(stderr, "l_setindex failed\n"); exit(7); }
    fill_lcg(vt, l, 4);
    if (l_lgs(g, vt, mA, vb, NULL) != NUM_OK) { fprintf(stderr, "l_lgs failed\n"); exit(7); }
    dumpvec("l_lgs", vt, l);
    fill_lcg(vt, l, 4);
    if (l_ugs(g, vt, mA, vb) != NUM_OK) { fprintf(stderr, "l_ugs failed\n"); exit(7); }
    dumpvec("l_ugs", vt, l);
    fill_lcg(vt, l, 4);
    if (l_lsor(g, vt, mA, vb, a3, NULL) != NUM_OK) { fprintf(stderr, "l_lsor failed\n"); exit(7); }
    dumpvec("l_lsor", vt, l);
    fill_lcg(vt, l, 4);
    if (l_usor(g, vt, mA, vb, a3, NULL) != NUM_OK) { fprintf(stderr, "l_usor failed\n"); exit(7); }
    dumpvec("l_usor", vt, l);
    fill_lcg(vt, l, 4);
    // ILU: what ILUPreProcess / ILUStep do (iter.cc:5444-5508): L = copy of A, l_ilubthdecomp (ugiter.cc:2252) with
    // the class's beta and no threshold, then l_luiter (:4444).  The decomposed values are dumped in the canonical entry order.
    if (o.smoother == "ilu") {
      MATDATA_DESC *mL = NULL;
      VEC_SCALAR beta;
      for (int i = 0; i < MAX_VEC_COMP; i++) beta[i] = o.beta;
      if (AllocMDFromMD(mg, l, l, mA, &mL)) { fprintf(stderr, "AllocMDFromMD failed\n"); exit(7); }
      if (dmatcopy(mg, l, l, ALL_VECTORS, mL, mA) != NUM_OK) { fprintf(stderr, "dmatcopy failed\n"); exit(7); }
      if (l_ilubthdecomp(g, mL, beta, NULL, NULL, NULL) != NUM_OK) { fprintf(stderr, "l_ilubthdecomp failed\n"); exit(7); }
      gpuls::FlatLevel fL;
      if (gpuls::FlattenFlags(mg, l, vx, fL) || gpuls::FlattenMatrix(mg, l, mL, fL)) { fprintf(stderr, "FlattenMatrix(L) failed\n"); exit(7); }
      D.f64(L("ilu/val", l), fL.val);
      if (l_luiter(g, vt, mL, vb) != NUM_OK) { fprintf(stderr, "l_luiter failed\n"); exit(7); }
      dumpvec("l_luiter", vt, l);
      fill_lcg(vt, l, 4);
      if (FreeMD(mg, l, l, mL)) { fprintf(stderr, "FreeMD failed\n"); exit(7); }
    }
    if (!o.lean) {
    // BLAS-2 (ALL_VECTORS, single level)
    dmatmul(mg, l, l, ALL_VECTORS, vt, mA, vx);       dumpvec("dmatmul", vt, l);
    dmatmul_add(mg, l, l, ALL_VECTORS, vt, mA, vb);   dumpvec("dmatmul_add", vt, l);
    dmatmul_minus(mg, l, l, ALL_VECTORS, vt, mA, vc); dumpvec("dmatmul_minus", vt, l);
    // BLAS-1 chain on t (each result depends on the previous one; inputs x,b,c stay fixed)
    fill_lcg(vt, l, 4);
    dcopy(mg, l, l, ALL_VECTORS, vt, vx);             dumpvec("dcopy", vt, l);
    dscal(mg, l, l, ALL_VECTORS, vt, 0.75);           dumpvec("dscal", vt, l);
    dscalx(mg, l, l, ALL_VECTORS, vt, a3);            dumpvec("dscalx", vt, l);
    dadd(mg, l, l, ALL_VECTORS, vt, vb);              dumpvec("dadd", vt, l);
    dsub(mg, l, l, ALL_VECTORS, vt, vc);              dumpvec("dsub", vt, l);
    dminusadd(mg, l, l, ALL_VECTORS, vt, vb);         dumpvec("dminusadd", vt, l);
    daxpy(mg, l, l, ALL_VECTORS, vt, -1.375, vx);     dumpvec("daxpy", vt, l);
    daxpyx(mg, l, l, ALL_VECTORS, vt, a3, vc);        dumpvec("daxpyx", vt, l);
    DOUBLE s; VEC_SCALAR sx;
    ddot(mg, l, l, ALL_VECTORS, vx, vb, &s);          D.scalar_d(L("ddot", l), s);
    dnrm2(mg, l, l, ALL_VECTORS, vx, &s);             D.scalar_d(L("dnrm2", l), s);
    ddotx(mg, l, l, ALL_VECTORS, vx, vb, sx);         D.rec(L("ddotx", l), 1, sx, BS, 8);
    dnrm2x(mg, l, l, ALL_VECTORS, vx, sx);            D.rec(L("dnrm2x", l), 1, sx, BS, 8);
    dset(mg, l, l, ALL_VECTORS, vt, 0.5);             dumpvec("dset", vt, l);
    }
    // l_jac and the smoother step of the configured class in defect-correction form
    fill_lcg(vt, l, 4);
    if (l_jac(g, vt, mA, vb) != NUM_OK) { fprintf(stderr, "l_jac failed\n"); exit(7); }
    dumpvec("l_jac", vt, l);
    if (l > 0) {
      fill_lcg(vt, l, 4);
      (*smooth->PreProcess)(smooth, l, vx, vb, mA, &bl, &result);
      if ((*smooth->Iter)(smooth, l, vt, vb, mA, &result)) { fprintf(stderr, "smoother failed\n"); exit(7); }
      dumpvec("smooth/t", vt, l); dumpvec("smooth/b", vb, l);
      (*smooth->PostProcess)(smooth, l, vx, vb, mA, &result);
      fill_lcg(vb, l, 2);
    }
  }
  D.rec("ops/a3", 1, a3, BS, 8);
  if (o.lean) return;
  // grid transfer: restrict x (fine) into c (coarse, pre-filled), prolong x (coarse) into t (fine)
  for (int l = 1; l <= top; l++) {
    fill_lcg(vx, l, 1); fill_lcg(vx, l - 1, 1); fill_lcg(vc, l - 1, 3); fill_lcg(vt, l, 4);
    // `to` and `from` are the same descriptor in Lmgc (b,b); use the same here: c on both levels
    fill_lcg(vc, l, 5);
    dumpvec("restrict/in_fine", vc, l); dumpvec("restrict/in_coarse", vc, l - 1);
    if ((o.imat ? RestrictByMatrix(GRID_ON_LEVEL(mg, l), vc, vc, a3) : StandardRestrict(GRID_ON_LEVEL(mg, l), vc, vc, a3)) != NUM_OK) { fprintf(stderr, "restrict failed\n"); exit(8); }
    dumpvec("restrict/out", vc, l - 1);
    if ((o.imat ? InterpolateCorrectionByMatrix(GRID_ON_LEVEL(mg, l), vt, vx, a3) : StandardInterpolateCorrection(GRID_ON_LEVEL(mg, l), vt, vx, a3)) != NUM_OK) { fprintf(stderr, "interpolate failed\n"); exit(8); }
    dumpvec("interpolate/in_coarse", vx, l - 1); dumpvec("interpolate/out", vt, l);
  }
  // surface-mode loops over all levels (matter on adaptive hierarchies)
  {
    int fr = FULLREFINELEVEL(mg);
    for (int l = 0; l <= top; l++) { fill_lcg(vx, l, 11); fill_lcg(vb, l, 12); fill_lcg(vt, l, 13); }
    for (int l = 0; l <= top; l++) { dumpvec("surf/in_x", vx, l); dumpvec("surf/in_b", vb, l); dumpvec("surf/in_t", vt, l); }
    dmatmul_minus(mg, fr, top, ON_SURFACE, vb, mA, vx);
    for (int l = 0; l <= top; l++) dumpvec("surf/dmatmul_minus", vb, l);
    VEC_SCALAR sx; DOUBLE s;
    dnrm2x(mg, fr, top, ON_SURFACE, vb, sx); D.rec("surf/dnrm2x", 1, sx, BS, 8);
    ddot(mg, fr, top, ON_SURFACE, vb, vx, &s); D.scalar_d("surf/ddot", s);
    dset(mg, fr, top, ON_SURFACE, vt, 2.5);
    for (int l = 0; l <= top; l++) dumpvec("surf/dset", vt, l);
    daxpy(mg, fr, top, ON_SURFACE, vt, 0.5, vx);
    for (int l = 0; l <= top; l++) dumpvec("surf/daxpy", vt, l);
  }
}

static void restore_problem(void)
{
  // rhs and sol back to the assembled state (assemble() also rewrites the matrix: same values)
  assemble();
  // algebraic levels: work space of the cycle only; zeroed so that every run starts from the same state
  for (int l = LO; l < 0; l++)
    for (VECTOR *v = FIRSTVECTOR(GRID_ON_LEVEL(mg, l)); v != NULL; v = SUCCVC(v))
      for (int i = 0; i < BS; i++) { VVALUE(v, VD_CMP_OF_TYPE(vx, VTYPE(v), i)) = 0.0; VVALUE(v, VD_CMP_OF_TYPE(vb, VTYPE(v), i)) = 0.0; VVALUE(v, VD_CMP_OF_TYPE(vc, VTYPE(v), i)) = 0.0; VVALUE(v, VD_CMP_OF_TYPE(vt, VTYPE(v), i)) = 0.0; }
}

// one Lmgc cycle through the numproc interface + the full `ls` solve
static void dump_solve(const Opt &o)
{
  int top = TOPLEVEL(mg);
  INT result = 0, bl = 0;
  restore_problem();
  NP_ITER *lmgc = (NP_ITER *)GetNumProcByName(mg, "lmgc", ITER_CLASS_NAME);
  // --- a single cycle on the raw right-hand side (c = 0 on entry)
  (*lmgc->PreProcess)(lmgc, top, vx, vb, mA, &bl, &result);
  dset(mg, LO, top, ALL_VECTORS, vc, 0.0);
  if ((*lmgc->Iter)(lmgc, top, vc, vb, mA, &result)) { fprintf(stderr, "Lmgc failed\n"); exit(9); }
  for (int l = LO; l <= top; l++) { dumpvec("lmgc/c", vc, l); dumpvec("lmgc/b", vb, l); }
  (*lmgc->PostProcess)(lmgc, top, vx, vb, mA, &result);
  // --- full solve, history after every iteration (maxit = cycles, limits unreachable)
  restore_problem();
  NP_LINEAR_SOLVER *ls = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, "mgs", LINEAR_SOLVER_CLASS_NAME);
  LRESULT lr; memset(&lr, 0, sizeof lr);
  (*ls->PreProcess)(ls, top, vx, vb, mA, &bl, &result);
  (*ls->Defect)(ls, top, vx, vb, mA, &result);
  (*ls->Residuum)(ls, bl, top, vx, vb, mA, &lr);
  D.rec("solve/first_defect", 1, lr.last_defect, BS, 8);
  for (int l = LO; l <= top; l++) dumpvec("solve/b_first", vb, l);
  (*ls->PostProcess)(ls, top, vx, vb, mA, &result);
  // iterate one cycle at a time to record the history: maxit=1 solver calls are NOT equivalent
  // (first_defect handling), so we call Solver once with maxit=cycles and read PCR-independent
  // results; the per-iteration history comes from re-running with increasing maxit on fresh data.
  std::vector<double> hist;
  for (int k = 1; k <= o.cycles; k++) {
    restore_problem();
    cmd("npinit mgs $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I lmgc $display no", k);
    (*ls->PreProcess)(ls, top, vx, vb, mA, &bl, &result);
    (*ls->Defect)(ls, top, vx, vb, mA, &result);
    (*ls->Residuum)(ls, bl, top, vx, vb, mA, &lr);
    VEC_SCALAR abslimit, red;
    for (int i = 0; i < MAX_VEC_COMP; i++) { abslimit[i] = 1e-30; red[i] = 1e-30; }
    if ((*ls->Solver)(ls, top, vx, vb, mA, abslimit, red, &lr)) { fprintf(stderr, "Solver failed\n"); exit(9); }
    for (int i = 0; i < BS; i++) hist.push_back(lr.last_defect[i]);
    (*ls->PostProcess)(ls, top, vx, vb, mA, &result);
    if (k == 1 || k == 2 || k == 5 || k == o.cycles) {
      char nm[64];
      for (int l = LO; l <= top; l++) {
        snprintf(nm, sizeof nm, "solve/x_after_%d", k); dumpvec(nm, vx, l);
        snprintf(nm, sizeof nm, "solve/b_after_%d", k); dumpvec(nm, vb, l);
      }
    }
  }
  D.f64("solve/history", hist);
  D.scalar_i("solve/cycles", o.cycles);
  printf("history:");
  for (size_t i = 0; i < hist.size(); i++) printf(" %.10e", hist[i]);
  printf("\n");
}

// Krylov accelerators of the reference around the same cycle: class `cg` (LinearSolver + CGUpdate,
// ls.cc:989) and class `bcgs` (BCGSSolver, ls.cc:1864), $I lmgc.  As in dump_solve, run k is a fresh solve with $m k.
static void dump_krylov(const Opt &o)
{
  int top = TOPLEVEL(mg);
  INT result = 0, bl = 0;
  cmd("npcreate kcg $c cg");
  cmd("npcreate kbcgs $c bcgs");
  const char *names[2] = {"cg", "bcgs"};
  const int K[2] = {6, 4};
  for (int w = 0; w < 2; w++) {
    std::vector<double> hist;
    std::vector<int32_t> nits;
    char key[64];
    for (int k = 1; k <= K[w]; k++) {
      restore_problem();
      cmd("npinit k%s $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I lmgc $display no", names[w], k);
      std::string npn = std::string("k") + names[w];
      NP_LINEAR_SOLVER *s = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, npn.c_str(), LINEAR_SOLVER_CLASS_NAME);
      if (!s) { fprintf(stderr, "numproc %s missing\n", npn.c_str()); exit(9); }
      LRESULT lr; memset(&lr, 0, sizeof lr);
      VEC_SCALAR abslimit, red;
      for (int i = 0; i < MAX_VEC_COMP; i++) { abslimit[i] = 1e-30; red[i] = 1e-30; }
      if ((*s->PreProcess)(s, top, vx, vb, mA, &bl, &result)) { fprintf(stderr, "%s PreProcess failed\n", names[w]); exit(9); }
      (*s->Defect)(s, top, vx, vb, mA, &result);
      (*s->Residuum)(s, bl, top, vx, vb, mA, &lr);
      if (k == 1) { snprintf(key, sizeof key, "%s/first_defect", names[w]); D.rec(key, 1, lr.last_defect, BS, 8); }
      if ((*s->Solver)(s, top, vx, vb, mA, abslimit, red, &lr)) { fprintf(stderr, "%s Solver failed\n", names[w]); exit(9); }
      for (int i = 0; i < BS; i++) hist.push_back(lr.last_defect[i]);
      nits.push_back(lr.number_of_linear_iterations);
      (*s->PostProcess)(s, top, vx, vb, mA, &result);
      if (k == 1 || k == 2 || k == K[w])
        for (int l = LO; l <= top; l++) {
          snprintf(key, sizeof key, "%s/x_after_%d", names[w], k); dumpvec(key, vx, l);
          snprintf(key, sizeof key, "%s/b_after_%d", names[w], k); dumpvec(key, vb, l);
        }
    }
    snprintf(key, sizeof key, "%s/history", names[w]); D.f64(key, hist);
    snprintf(key, sizeof key, "%s/iterations", names[w]); D.i32(key, nits);
    snprintf(key, sizeof key, "%s/K", names[w]); D.scalar_i(key, K[w]);
    printf("%s history:", names[w]);
    for (size_t i = 0; i < hist.size(); i++) printf(" %.6e", hist[i]);
    printf("\n");
  }
}

// Rendezvous of concurrently started replicas (bench.py runs one per host core, UG being single-threaded): every
// replica drops a file when its hierarchy is built and waits for all the others, so the timed solves overlap.
static void replica_barrier(const Opt &o)
{
  if (o.barrier_dir.empty() || o.barrier_n <= 1) return;
  char nm[512];
  snprintf(nm, sizeof nm, "%s/ready.%d", o.barrier_dir.c_str(), o.barrier_id);
  FILE *f = fopen(nm, "w"); if (f) fclose(f);
  for (int tries = 0; tries < 600000; tries++) {
    int have = 0;
    for (int i = 0; i < o.barrier_n; i++) {
      snprintf(nm, sizeof nm, "%s/ready.%d", o.barrier_dir.c_str(), i);
      FILE *g = fopen(nm, "r"); if (g) { fclose(g); have++; }
    }
    if (have >= o.barrier_n) return;
    struct timespec ts = {0, 2000000}; nanosleep(&ts, NULL);
  }
}

// CPU baseline: time the reference's own solver (1 core per process; UG is single-threaded)
static void time_reference(const Opt &o)
{
  int top = TOPLEVEL(mg);
  INT result = 0, bl = 0;
  long n = NVEC(GRID_ON_LEVEL(mg, top));
  restore_problem();
  NP_LINEAR_SOLVER *ls = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, "mgs", LINEAR_SOLVER_CLASS_NAME);
  cmd("npinit mgs $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I lmgc $display no", o.cycles);
  LRESULT lr; memset(&lr, 0, sizeof lr);
  if ((*ls->PreProcess)(ls, top, vx, vb, mA, &bl, &result)) { fprintf(stderr, "PreProcess of the reference's numprocs failed\n"); exit(9); }
  (*ls->Defect)(ls, top, vx, vb, mA, &result);
  (*ls->Residuum)(ls, bl, top, vx, vb, mA, &lr);
  VEC_SCALAR abslimit, red;
  for (int i = 0; i < MAX_VEC_COMP; i++) { abslimit[i] = 1e-30; red[i] = 1e-30; }
  replica_barrier(o);
  double t0 = now();
  (*ls->Solver)(ls, top, vx, vb, mA, abslimit, red, &lr);
  double t1 = now();
  (*ls->PostProcess)(ls, top, vx, vb, mA, &result);
  int its = lr.number_of_linear_iterations ? lr.number_of_linear_iterations : o.cycles;
  double tcyc = (t1 - t0) / its;
  // kernel timings
  double best_mm = 1e30, best_dot = 1e30;
  for (int r = 0; r < o.reps; r++) {
    double a = now(); dmatmul_minus(mg, top, top, ALL_VECTORS, vb, mA, vx); double b = now();
    DOUBLE s; ddot(mg, top, top, ALL_VECTORS, vb, vx, &s); double c = now();
    if (b - a < best_mm) best_mm = b - a;
    if (c - b < best_dot) best_dot = c - b;
  }
  long nnz = 0;
  for (VECTOR *v = FIRSTVECTOR(GRID_ON_LEVEL(mg, top)); v; v = SUCCVC(v)) for (MATRIX *m = VSTART(v); m; m = MNEXT(m)) nnz++;
  printf("{\"kind\": \"reference\", \"cores\": 1, \"dim\": %d, \"bs\": %d, \"levels\": %d, \"unknowns\": %ld, \"nnz\": %ld, "
         "\"cycles\": %d, \"s_per_cycle\": %.6e, \"vcycle_unknowns_per_s\": %.6e, \"dmatmul_minus_s\": %.6e, \"ddot_s\": %.6e, "
         "\"last_defect\": %.10e}\n",
         DIM, BS, top + 1, n * BS, nnz, its, tcyc, n * BS / tcyc, best_mm, best_dot, lr.last_defect[0]);
}

// Galerkin coarse-grid operators: per level what `npcheck $G` does (np/algebra/npcheck.cc:375-379) -- dmatset(coarse, 0),
// AssembleGalerkinByMatrix(fine grid, A, 0) (np/algebra/transgrid.cc:1575) on the stored interpolation matrices -- cascaded from the
// top level down, so level l-1 is built from the Galerkin matrix of level l.  Connections the product needs and the coarse pattern
// lacks are created by the reference (CreateExtraConnection: second place of both row lists); the dump holds the resulting
// pattern and values of every coarse level in the canonical entry order.  Last record group of a dump: the matrices stay changed.
static void dump_galerkin(const Opt &o)
{
  (void)o;
  int top = TOPLEVEL(mg);
  restore_problem();
  for (int l = top; l >= 1; l--) {
    if (dmatset(mg, l - 1, l - 1, ALL_VECTORS, mA, 0.0) != NUM_OK) { fprintf(stderr, "dmatset failed\n"); exit(11); }
    if (AssembleGalerkinByMatrix(GRID_ON_LEVEL(mg, l), mA, 0) != NUM_OK) { fprintf(stderr, "AssembleGalerkinByMatrix failed\n"); exit(11); }
    gpuls::FlatLevel f;
    if (gpuls::FlattenFlags(mg, l - 1, vx, f) || gpuls::FlattenMatrix(mg, l - 1, mA, f)) { fprintf(stderr, "FlattenMatrix(galerkin) failed\n"); exit(11); }
    D.i32(L("galerkin/rowptr", l - 1), f.rowptr); D.i32(L("galerkin/col", l - 1), f.col); D.f64(L("galerkin/val", l - 1), f.val);
  }
}

#ifdef WITH_GPULS
static int run_gpu(const Opt &o);
#endif

int main(int argc, char **argv)
{
  Opt o;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto nxt = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return std::string(argv[++i]); };
    if (a == "--grid") o.grid = nxt(); else if (a == "--bs") o.bs = atoi(nxt().c_str());
    else if (a == "--refine") o.refine = atoi(nxt().c_str()); else if (a == "--adapt") o.adapt = atoi(nxt().c_str());
    else if (a == "--nu1") o.nu1 = atoi(nxt().c_str()); else if (a == "--nu2") o.nu2 = atoi(nxt().c_str());
    else if (a == "--gamma") o.gamma = atoi(nxt().c_str()); else if (a == "--cycles") o.cycles = atoi(nxt().c_str());
    else if (a == "--reps") o.reps = atoi(nxt().c_str());
    else if (a == "--barrier") { o.barrier_dir = nxt(); o.barrier_n = atoi(nxt().c_str()); o.barrier_id = atoi(nxt().c_str()); }
    else if (a == "--damp") o.damp = atof(nxt().c_str()); else if (a == "--dump") o.dump = nxt();
    else if (a == "--ops") o.ops = true; else if (a == "--solve") o.solve = true; else if (a == "--time") o.timeit = true;
    else if (a == "--verbose") o.quiet = false; else if (a == "--gpu") o.gpu = nxt();
    else if (a == "--smoother") o.smoother = nxt(); else if (a == "--baselevel") o.baselevel = atoi(nxt().c_str());
    else if (a == "--lean") o.lean = true; else if (a == "--imat") o.imat = true;
    else if (a == "--beta") o.beta = atof(nxt().c_str());
    else if (a == "--galerkin") o.galerkin = true;
    else if (a == "--assemble") { o.assemble = true; o.elems = true; }
    else if (a == "--savedata") o.savedata = nxt();      // prefix of the data files the reference's SaveData writes (np/udm/data_io.cc:650)
    else if (a == "--nokrylov") o.nokrylov = true;
    else if (a == "--amg") { o.amg_class = nxt(); o.amg_init = nxt(); }
    else if (a == "--gpuamg") o.gpuamg = nxt();
    else if (a == "--levelopt") o.levelopt = true;
    else if (a == "--hooks") o.hooks = true; else if (a == "--transferD") o.transferD = true;
    else if (a == "--collapse") o.collapse = true; else if (a == "--refine2") o.refine2 = atoi(nxt().c_str());
    else if (a == "--elems") o.elems = true;             // dump the elements (corner rows, fathers): input of the element partition (ug_b200/partition.py)       // --gpu: only the ls/lmgc mixes (bench.py's equal-size line)
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
  }
  int ac = 1; char *av0 = argv[0]; char **av = &av0;
  // keep UG's banner noise out of stdout: results are printed by this driver only
  FILE *saved = stdout;
  (void)saved;
  signal(SIGSEGV, on_crash); signal(SIGABRT, on_crash);
  if (InitUg(&ac, &av)) { fprintf(stderr, "InitUg failed\n"); return 1; }
#ifdef WITH_GPULS
  if (!o.gpu.empty()) { if (gpuls::LoadDeviceLibrary(o.gpu.c_str())) { fprintf(stderr, "cannot load %s\n", o.gpu.c_str()); return 1; } if (InitGpuLS()) { fprintf(stderr, "InitGpuLS failed\n"); return 1; } }
#endif
  double t0 = now();
  build_hierarchy(o);
  double t1 = now();
  int top = TOPLEVEL(mg);
  printf("hierarchy: dim=%d grid=%s bs=%d levels=%d fullrefinelevel=%d n=[", DIM, o.grid.c_str(), BS, top + 1, (int)FULLREFINELEVEL(mg));
  for (int l = 0; l <= top; l++) printf("%s%d", l ? "," : "", (int)NVEC(GRID_ON_LEVEL(mg, l)));
  printf("] build_s=%.2f\n", t1 - t0);
  if (o.smoother != "jac" && o.smoother != "gs" && o.smoother != "sgs" && o.smoother != "sor" && o.smoother != "ilu") { fprintf(stderr, "unknown smoother %s\n", o.smoother.c_str()); return 1; }
  if (o.galerkin && !o.imat) { fprintf(stderr, "--galerkin needs --imat (the stored interpolation matrices)\n"); return 1; }
  if (o.imat)     // the interpolation matrices the $M mode works on (transgrid.cc:2363); the format reserves them ($I)
    for (int l = 1; l <= top; l++)
      if (CreateStandardNodeRestProl(GRID_ON_LEVEL(mg, l), BS) != NUM_OK) { fprintf(stderr, "CreateStandardNodeRestProl failed\n"); return 1; }
  make_numprocs(o, "", o.smoother.c_str(), "lmgc", "transfer", "ls", o.cycles);
  std::vector<gpuls::FlatLevel> fl;
  if (!o.dump.empty() && !o.amg_class.empty()) {
    // the algebraic levels must exist while the dump is written: one PreProcess / PostProcess bracket of the transfer builds them, and
    // the AMG numproc's $hold keeps them (later brackets keep coarsening and interpolation and recompute the matrices, amgtransfer.cc:1004)
    if (o.amg_init.find("$hold") == std::string::npos || o.ops || o.galerkin || o.assemble || !o.savedata.empty()) { fprintf(stderr, "--amg dumps need $hold and hold the hierarchy and the solve records only\n"); return 1; }
    NP_TRANSFER *t = (NP_TRANSFER *)GetNumProcByName(mg, "transfer", TRANSFER_CLASS_NAME);
    INT fl0 = 0, result = 0;
    if (!t || (*t->PreProcess)(t, &fl0, top, vx, vb, mA, &result) || (*t->PostProcess)(t, &fl0, top, vx, vb, mA, &result)) { fprintf(stderr, "AMG setup failed\n"); return 1; }
    LO = BOTTOMLEVEL(mg);
    printf("algebraic levels: bottom=%d n=[", LO);
    for (int l = LO; l < 0; l++) printf("%s%d", l > LO ? "," : "", (int)NVEC(GRID_ON_LEVEL(mg, l)));
    printf("]\n");
    restore_problem();
  }
  if (!o.dump.empty()) {
    D.open(o.dump.c_str());
    dump_hierarchy(o, fl);
    if (o.ops) dump_ops(o);
    if (o.solve) dump_solve(o);
    if (o.solve && !o.lean) dump_krylov(o);
    if (o.assemble) dump_assemble(o);
    if (!o.savedata.empty()) dump_savedata(o);
    if (o.galerkin) dump_galerkin(o);
    D.close();
  }
  if (o.timeit) time_reference(o);
#ifdef WITH_GPULS
  if (!o.gpu.empty()) return run_gpu(o);
#endif
  return 0;
}

#ifdef WITH_GPULS
// Drop-in test: identical numproc script with the gpuls classes; compare with the CPU classes.
static double maxrel(const std::vector<double> &a, const std::vector<double> &b)
{
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); i++) { num = fmax(num, fabs(a[i] - b[i])); den = fmax(den, fabs(a[i])); }
  return den > 0 ? num / den : num;
}

static int run_gpu(const Opt &o)
{
  int top = TOPLEVEL(mg);
  INT result = 0, bl = 0;
  VEC_SCALAR abslimit, red;
  for (int i = 0; i < MAX_VEC_COMP; i++) { abslimit[i] = 1e-30; red[i] = 1e-30; }
  int fails = 0;
  // 1. CPU reference run
  restore_problem();
  cmd("npinit mgs $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I lmgc $display no", o.cycles);
  NP_LINEAR_SOLVER *ls = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, "mgs", LINEAR_SOLVER_CLASS_NAME);
  LRESULT lr; memset(&lr, 0, sizeof lr);
  double bscale = 1e-300;
  { std::vector<double> b0 = gather(vb, top); for (size_t i = 0; i < b0.size(); i++) bscale = fmax(bscale, fabs(b0[i])); }
  const double cp0 = now();
  if ((*ls->PreProcess)(ls, top, vx, vb, mA, &bl, &result)) { printf("FAIL reference: PreProcess of the CPU numprocs\n"); return 1; }
  const double cpre = now() - cp0;      // PreProcess: with --amg this is where the algebraic levels are built
  (*ls->Defect)(ls, top, vx, vb, mA, &result);
  (*ls->Residuum)(ls, bl, top, vx, vb, mA, &lr);
  double c0 = now();
  (*ls->Solver)(ls, top, vx, vb, mA, abslimit, red, &lr);
  double c1 = now();
  (*ls->PostProcess)(ls, top, vx, vb, mA, &result);
  std::vector<std::vector<double> > xr(top + 1), br(top + 1);
  for (int l = 0; l <= top; l++) { xr[l] = gather(vx, l); br[l] = gather(vb, l); }
  LRESULT lr_cpu = lr;

  // 2. every gpuls configuration: (a) all four GPU classes, device base solver; (b) GPU classes with the
  //    CPU base solver numproc; (c) CPU ls + gpulmgc; (d) CPU ls + CPU lmgc + gpujac + gputransfer
  struct Cfg { const char *name, *jac, *lmgc, *transfer, *ls; const char *extra; double tol; };
  const Cfg cfgs[] = {
    {"gpuls+gpulmgc+gpujac+gputransfer (host base solver)", "gpujac", "gpulmgc", "gputransfer", "gpuls", "", 0.0},
    {"gpuls+gpulmgc, device base solver", "gpujac", "gpulmgc", "gputransfer", "gpuls", " $devbase", 1e-12},
    {"ls+gpulmgc", "gpujac", "gpulmgc", "gputransfer", "ls", "", 0.0},
    {"ls+lmgc+gpujac+gputransfer", "gpujac", "lmgc", "gputransfer", "ls", "", 0.0},
  };
  int k = 0;
  for (const Cfg &c : cfgs) {
    char pfx[16]; snprintf(pfx, sizeof pfx, "g%d", k++);
    // algebraic levels on the device only: no host vectors there, so only the device-resident solve with the device base solver applies
    if (!o.gpuamg.empty() && strstr(c.extra, "devbase") == NULL) continue;
    make_numprocs(o, pfx, (std::string("gpu") + o.smoother).c_str(), c.lmgc, c.transfer, c.ls, o.cycles);
    if (c.extra[0]) cmd("npinit %slmgc $S %ssmooth %ssmooth %sbasesolver $T %stransfer $n1 %d $n2 %d $g %d $b %d%s", pfx, pfx, pfx, pfx, pfx, o.nu1, o.nu2, o.gamma, o.baselevel, c.extra);
    restore_problem();
    std::string name = std::string(pfx) + "mgs";
    NP_LINEAR_SOLVER *g = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, name.c_str(), LINEAR_SOLVER_CLASS_NAME);
    memset(&lr, 0, sizeof lr);
    const double gp0 = now();
    if ((*g->PreProcess)(g, top, vx, vb, mA, &bl, &result)) { printf("FAIL %s: PreProcess\n", c.name); fails++; continue; }
    const double gpre = now() - gp0;      // flattening + upload (+ the algebraic levels: the reference's numproc on the host, or $gpuamg)
    (*g->Defect)(g, top, vx, vb, mA, &result);
    (*g->Residuum)(g, bl, top, vx, vb, mA, &lr);
    double g0 = now();
    if ((*g->Solver)(g, top, vx, vb, mA, abslimit, red, &lr)) { printf("FAIL %s: Solver\n", c.name); fails++; continue; }
    double g1 = now();
    (*g->PostProcess)(g, top, vx, vb, mA, &result);
    double ex = 0, eb = 0;
    for (int l = 0; l <= top; l++) { ex = fmax(ex, maxrel(xr[l], gather(vx, l))); eb = fmax(eb, maxrel(br[l], gather(vb, l))); }
    if (o.levelopt) {
      // results to rounding: a defect that an exact base solve leaves at rounding level has no scale of its own -- defects are compared on
      // the scale of the first defect (as for the Krylov classes below)
      eb = 0;
      for (int l = 0; l <= top; l++) { std::vector<double> gb = gather(vb, l); for (size_t i = 0; i < gb.size(); i++) eb = fmax(eb, fabs(gb[i] - br[l][i]) / bscale); }
    }
    double ed = 0;
    for (int i = 0; i < BS; i++) ed = fmax(ed, fabs(lr.last_defect[i] - lr_cpu.last_defect[i]) / lr_cpu.last_defect[i]);
    // transfer $L: the two scalars of MinimizeLevel are parallel sums on the device -- agreement to rounding, like the Krylov classes
    const double tol = o.levelopt ? fmax(c.tol, 1e-11) : c.tol;
    bool ok = ex <= tol && eb <= tol && ed <= (o.levelopt ? 1e-11 : 1e-12) && lr.number_of_linear_iterations == lr_cpu.number_of_linear_iterations;
    printf("%s %s: its=%d last_defect=%.10e (cpu %.10e) relerr x=%.3e b=%.3e defect=%.3e  t_gpu=%.4fs t_cpu=%.4fs  pre_gpu=%.4fs pre_cpu=%.4fs\n", ok ? "PASS" : "FAIL", c.name,
           (int)lr.number_of_linear_iterations, lr.last_defect[0], lr_cpu.last_defect[0], ex, eb, ed, g1 - g0, c1 - c0, gpre, cpre);
    if (!ok) fails++;
  }
  // 2b. the nested-iteration hooks of NP_TRANSFER (transfer.h:79-166) on seeded vectors: class transfer vs class gputransfer
  if (o.hooks) {
    NP_TRANSFER *tc = (NP_TRANSFER *)GetNumProcByName(mg, "transfer", TRANSFER_CLASS_NAME), *tg = (NP_TRANSFER *)GetNumProcByName(mg, "g0transfer", TRANSFER_CLASS_NAME);
    std::vector<std::vector<double> > ref[2];
    bool okh = tc && tg && tg->InterpolateNewVectors && tg->ProjectSolution;
    for (int side = 0; okh && side < 2; side++) {
      NP_TRANSFER *t = side ? tg : tc;
      for (int l = 0; l <= top; l++) fill_lcg(vx, l, 21);
      // mark every vector of the levels above 0 as new, as a grid adaption would: the hook then interpolates all of them
      for (int l = 1; l <= top; l++) for (VECTOR *v = FIRSTVECTOR(GRID_ON_LEVEL(mg, l)); v != NULL; v = SUCCVC(v)) SETVNEW(v, 1);
      if ((*t->InterpolateNewVectors)(t, 0, top, vx, &result)) okh = false;
      std::vector<std::vector<double> > got;
      for (int l = 0; l <= top; l++) got.push_back(gather(vx, l));
      for (int l = 0; l <= top; l++) fill_lcg(vx, l, 22);
      if ((*t->ProjectSolution)(t, 0, top, vx, &result)) okh = false;
      for (int l = 0; l <= top; l++) got.push_back(gather(vx, l));
      if (side == 0) ref[0] = got; else ref[1] = got;
    }
    okh = okh && ref[0] == ref[1];
    printf("%s hooks: InterpolateNewVectors / ProjectSolution of gputransfer vs transfer, %d levels, bitwise\n", okh ? "PASS" : "FAIL", top + 1);
    if (!okh) fails++;
    restore_problem();
  }
  // 3. Krylov accelerators: the reference's `cg` / `bcgs` around its own lmgc against gpucg / gpubcgs around gpulmgc
  //    (device-resident).  Step lengths come from parallel sums on the device: agreement to rounding (1e-9), not bitwise.
  if (!o.nokrylov) {
    const char *cpu_cls[2] = {"cg", "bcgs"}, *gpu_cls[2] = {"gpucg", "gpubcgs"};
    const int its[2] = {6, 4};
    for (int w = 0; w < 2; w++) {
      std::vector<std::vector<double> > xs(top + 1), bs_(top + 1);
      LRESULT lrs[2];
      double tm[2] = {0, 0};
      bool okrun = true;
      for (int side = 0; side < 2; side++) {
        char nm[32]; snprintf(nm, sizeof nm, "kd%d%d", w, side);
        cmd("npcreate %s $c %s", nm, side ? gpu_cls[w] : cpu_cls[w]);
        cmd("npinit %s $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I %slmgc $display no", nm, its[w], side ? "g0" : "");
        restore_problem();
        NP_LINEAR_SOLVER *g = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, nm, LINEAR_SOLVER_CLASS_NAME);
        memset(&lr, 0, sizeof lr);
        if (!g || (*g->PreProcess)(g, top, vx, vb, mA, &bl, &result)) { printf("FAIL %s: PreProcess\n", side ? gpu_cls[w] : cpu_cls[w]); okrun = false; break; }
        (*g->Defect)(g, top, vx, vb, mA, &result);
        (*g->Residuum)(g, bl, top, vx, vb, mA, &lr);
        double g0 = now();
        if ((*g->Solver)(g, top, vx, vb, mA, abslimit, red, &lr)) { printf("FAIL %s: Solver\n", side ? gpu_cls[w] : cpu_cls[w]); okrun = false; break; }
        tm[side] = now() - g0;
        (*g->PostProcess)(g, top, vx, vb, mA, &result);
        lrs[side] = lr;
        if (side == 0) for (int l = 0; l <= top; l++) { xs[l] = gather(vx, l); bs_[l] = gather(vb, l); }
      }
      if (!okrun) { fails++; continue; }
      double ex = 0, eb = 0, bscale = 0;
      for (int l = 0; l <= top; l++) {
        ex = fmax(ex, maxrel(xs[l], gather(vx, l)));
        std::vector<double> gb = gather(vb, l);
        for (size_t i = 0; i < gb.size(); i++) eb = fmax(eb, fabs(gb[i] - bs_[l][i]));
      }
      for (int i = 0; i < BS; i++) bscale = fmax(bscale, lrs[0].first_defect[i]);
      eb /= bscale > 0 ? bscale : 1.0;
      bool ok = ex <= 1e-9 && eb <= 1e-9 && lrs[0].number_of_linear_iterations == lrs[1].number_of_linear_iterations;
      printf("%s %s vs %s: its=%d/%d last_defect=%.6e (cpu %.6e) relerr x=%.3e b=%.3e (of the first defect)  t_gpu=%.4fs t_cpu=%.4fs\n", ok ? "PASS" : "FAIL",
             gpu_cls[w], cpu_cls[w], (int)lrs[1].number_of_linear_iterations, (int)lrs[0].number_of_linear_iterations, lrs[1].last_defect[0], lrs[0].last_defect[0],
             ex, eb, tm[1], tm[0]);
      if (!ok) fails++;
    }
  }
  // 4. assemble.gpufe against the reference's NP_LOCAL_ASSEMBLE loop with the same element kernel (class `fe` above): every MVALUE,
  //    the right-hand side, x (Dirichlet values) and VECSKIP of every level bit for bit; then a gpuls solve INSIDE gpufe's bracket
  //    (the matrix never crosses PCIe as values) against the CPU solve of the same assembled problem
  if (o.assemble) {
    struct Snap { std::vector<double> val, b, x; std::vector<uint32_t> skip; };
    auto snap = [&](std::vector<Snap> &out) {
      out.assign(top + 1, Snap());
      for (int l = 0; l <= top; l++) {
        gpuls::FlatLevel f;
        if (gpuls::FlattenFlags(mg, l, vx, f) || gpuls::FlattenMatrix(mg, l, mA, f)) { fprintf(stderr, "flatten failed\n"); exit(12); }
        out[l].val = f.val; out[l].skip = f.skip; out[l].b = gather(vb, l); out[l].x = gather(vx, l);
      }
    };
    std::vector<Snap> ref, got;
    run_fe_assemble("fe");
    snap(ref);
    cmd("npinit mgs $A MAT $x sol $b rhs $m %d $red 1e-30 $abslimit 1e-30 $I lmgc $display no", o.cycles);
    memset(&lr, 0, sizeof lr);
    (*ls->PreProcess)(ls, top, vx, vb, mA, &bl, &result);
    (*ls->Defect)(ls, top, vx, vb, mA, &result);
    (*ls->Residuum)(ls, bl, top, vx, vb, mA, &lr);
    (*ls->Solver)(ls, top, vx, vb, mA, abslimit, red, &lr);
    (*ls->PostProcess)(ls, top, vx, vb, mA, &result);
    std::vector<std::vector<double> > xs(top + 1);
    for (int l = 0; l <= top; l++) xs[l] = gather(vx, l);
    LRESULT lr_ref = lr;
    // wipe what the device run has to produce
    for (int l = 0; l <= top; l++) {
      dmatset(mg, l, l, ALL_VECTORS, mA, -7.0); dset(mg, l, l, ALL_VECTORS, vb, -7.0);
      for (VECTOR *v = FIRSTVECTOR(GRID_ON_LEVEL(mg, l)); v; v = SUCCVC(v)) VECSKIP(v) = 0;
    }
    gpuls::SetFEData(fe_coef, fe_dirichlet);
    NP_ASSEMBLE *gass = run_fe_assemble("gpufe", true);
    snap(got);
    size_t bad = 0, cnt = 0;
    for (int l = 0; l <= top; l++) {
      for (size_t i = 0; i < ref[l].val.size(); i++, cnt++) if (memcmp(&ref[l].val[i], &got[l].val[i], 8)) bad++;
      for (size_t i = 0; i < ref[l].b.size(); i++, cnt++) if (memcmp(&ref[l].b[i], &got[l].b[i], 8) || memcmp(&ref[l].x[i], &got[l].x[i], 8)) bad++;
      for (size_t i = 0; i < ref[l].skip.size(); i++, cnt++) if (ref[l].skip[i] != got[l].skip[i]) bad++;
    }
    printf("%s gpufe vs fe (NP_LOCAL_ASSEMBLE): %zu of %zu values differ (matrix, rhs, sol, VECSKIP on %d levels)\n", bad ? "FAIL" : "PASS", bad, cnt, top + 1);
    if (bad) fails++;
    NP_LINEAR_SOLVER *g = (NP_LINEAR_SOLVER *)GetNumProcByName(mg, "g1mgs", LINEAR_SOLVER_CLASS_NAME);
    memset(&lr, 0, sizeof lr);
    if (!g || (*g->PreProcess)(g, top, vx, vb, mA, &bl, &result)) { printf("FAIL gpuls inside the gpufe bracket: PreProcess\n"); fails++; }
    else {
      (*g->Defect)(g, top, vx, vb, mA, &result);
      (*g->Residuum)(g, bl, top, vx, vb, mA, &lr);
      int rc = (*g->Solver)(g, top, vx, vb, mA, abslimit, red, &lr);
      (*g->PostProcess)(g, top, vx, vb, mA, &result);
      double ex = 0;
      for (int l = 0; l <= top; l++) ex = fmax(ex, maxrel(xs[l], gather(vx, l)));
      double ed = fabs(lr.last_defect[0] - lr_ref.last_defect[0]) / lr_ref.last_defect[0];
      bool ok = !rc && ex <= 1e-12 && ed <= 1e-12 && lr.number_of_linear_iterations == lr_ref.number_of_linear_iterations;
      printf("%s gpuls+gpulmgc inside the gpufe bracket (matrix assembled on the device, no value upload): its=%d last_defect=%.10e (cpu %.10e) relerr x=%.3e\n",
             ok ? "PASS" : "FAIL", (int)lr.number_of_linear_iterations, lr.last_defect[0], lr_ref.last_defect[0], ex);
      if (!ok) fails++;
    }
    // savedata / loaddata on the device mirror against the reference's SaveData (np/udm/data_io.cc:650): (a) the device copies of sol and
    // rhs -> file, bytes equal to the file the reference writes from the VVALUEs; (b) a file the reference wrote from OTHER values -> device
    // vectors (gpuls::LoadData) -> file again (gpuls::SaveData): bytes equal to the file that was loaded
    {
      VECDATA_DESC *vds[2] = {vx, vb};
      EVALUES *ev[2] = {NULL, NULL}; EVECTOR *evec[2] = {NULL, NULL};
      char base[64]; snprintf(base, sizeof base, "/tmp/ugsd_%d", (int)getpid());
      std::string ref1 = std::string(base) + "_ref1", dev1 = std::string(base) + "_dev1", ref2 = std::string(base) + "_ref2", dev2 = std::string(base) + "_dev2";
      auto bytes = [](const std::string &p) { std::vector<char> b; FILE *f = fopen(p.c_str(), "rb"); if (f) { for (int c; (c = fgetc(f)) != EOF;) b.push_back((char)c); fclose(f); } return b; };
      const MEM heap_size_field = MGHEAP(mg)->size;
      int bad_io = 0;
      // after the solve above the VVALUEs of sol and rhs equal their device copies (gpuls downloads both at the end of Solver)
      std::vector<std::vector<double> > cx(top + 1), cb(top + 1);
      for (int l = 0; l <= top; l++) { cx[l] = gather(vx, l); cb[l] = gather(vb, l); }
      for (const char *type : {"bin", "asc"}) {
        for (int l = 0; l <= top; l++) { gpuls::ScatterVector(mg, l, vx, BS, cx[l].data()); gpuls::ScatterVector(mg, l, vb, BS, cb[l].data()); }
        MGHEAP(mg)->size = MGHEAP(mg)->used + (MEM)(1u << 30);
        int e1 = SaveData(mg, (char *)ref1.c_str(), 1, 1, (char *)type, -1, 0.0, 0.0, 0.0, 2, vds, ev, evec, NULL);
        for (int l = 0; l <= top; l++) { fill_lcg(vx, l, 21); fill_lcg(vb, l, 22); }
        int e2 = SaveData(mg, (char *)ref2.c_str(), 1, 1, (char *)type, -1, 0.0, 0.0, 0.0, 2, vds, ev, evec, NULL);
        MGHEAP(mg)->size = heap_size_field;
        int e3 = gpuls::SaveData(mg, dev1.c_str(), type, -1, 0.0, 0.0, 0.0, 2, vds);
        std::string sfx = std::string(".ug.data.") + type;
        std::vector<char> r1 = bytes(ref1 + sfx), d1 = bytes(dev1 + sfx), r2 = bytes(ref2 + sfx);
        bool ok1 = !e1 && !e3 && !r1.empty() && r1 == d1;
        printf("%s savedata %s from the device mirror: %zu bytes, %s the reference's file\n", ok1 ? "PASS" : "FAIL", type, d1.size(), ok1 ? "identical to" : "DIFFERENT from");
        bool ok2 = false;
        if (!e2 && !gpuls::LoadData(mg, ref2.c_str(), type, -1, 2, vds) && !gpuls::SaveData(mg, dev2.c_str(), type, -1, 0.0, 0.0, 0.0, 2, vds)) {
          std::vector<char> d2 = bytes(dev2 + sfx);
          ok2 = !r2.empty() && r2 == d2 && r2 != r1;
        }
        printf("%s loaddata %s into the device mirror, saved again: %s the loaded file\n", ok2 ? "PASS" : "FAIL", type, ok2 ? "identical to" : "DIFFERENT from");
        if (!ok1) bad_io++;
        if (!ok2) bad_io++;
        // the device copies go back to the values of (a) for the next mode
        if (strcmp(type, "bin") == 0 && gpuls::LoadData(mg, ref1.c_str(), "bin", -1, 2, vds)) bad_io++;
        for (const std::string &p : {ref1, dev1, ref2, dev2}) remove((p + sfx).c_str());
      }
      fails += bad_io;
    }
    (*gass->PostProcess)(gass, top, vx, vb, mA, &result);
    restore_problem();
  }
  printf("gpuls drop-in: %d failure(s)\n", fails);
  return fails ? 10 : 0;
}
#endif
