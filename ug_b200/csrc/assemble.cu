// assemble.cu -- element-loop assembly on the device: what NP_LOCAL_ASSEMBLE does on one level,
//   np/procs/assemble.cc:657-706 LocalAssemble   dset(b, 0), dmatset(A, 0), CLEAR_VECSKIP, then for the elements in list order
//                                                 GetElementVVMPtrs (np/udm/disctools.cc:1113), AssembleLocal, `*rptr += def`,
//                                                 `*mptr += mat`, SetElementDirichletFlags (:1763) on boundary elements,
//   np/procs/assemble.cc:624 NPLocalAssemblePostMatrix -> AssembleDirichletBoundary (disctools.cc:1837): per component with its VECSKIP
//                                                 bit set  b = x, the row of the diagonal block = unit row, the row of every other block = 0,
// with the element kernel (AssembleLocal -- application code in UG) built in: P1 / Q1 diffusion with one coefficient per element or
// isotropic linear elasticity, simplices with the centroid rule, tensor elements with 2-point Gauss, local matrix summed over the
// quadrature points before it is added to the global one.
//
// The reference SCATTERS element by element.  Here one thread per matrix ROW GATHERS: the (element, corner) incidences of its vector
// in element-list order (stable radix sort of the corner lists by row), for each the element's geometry and its local row, added to
// the row's entries in that order -- every stored value receives the reference's terms in the reference's order, no atomics:
// bit-identical values.  The element geometry is recomputed once per corner (4 x for tetrahedra, 8 x for hexahedra): arithmetic is
// free next to the gathers of a setup kernel.
#include <cub/device/device_radix_sort.cuh>       // before uggpu_internal.h (its SLICE macro)
#include <cub/device/device_scan.cuh>

#include "uggpu_internal.h"

#include <cmath>

struct FeParam {
  int problem;               // UGGPU_FE_POISSON / UGGPU_FE_ELASTICITY
  double lam, mu;
  double source[UGGPU_MAX_BS];
};

template <int DIM>
__device__ __forceinline__ double fe_det_inv(const double (&J)[DIM][DIM], double (&Ji)[DIM][DIM])
{
  if constexpr (DIM == 2) {
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    Ji[0][0] = J[1][1] / det; Ji[0][1] = -J[0][1] / det; Ji[1][0] = -J[1][0] / det; Ji[1][1] = J[0][0] / det;
    return det;
  } else {
    const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                       J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    Ji[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
    Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; Ji[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det;
    Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    Ji[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
    Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    return det;
  }
}

// quadrature point q of an element with NC corners X: weight * |det J|, shape values N, gradients G (global coordinates)
template <int DIM, int NC>
__device__ __forceinline__ void fe_qp(const double (&X)[NC][DIM], int q, double &w, double (&N)[NC], double (&G)[NC][DIM])
{
  double J[DIM][DIM], Ji[DIM][DIM];
  if constexpr (NC == DIM + 1) {
    // barycentric coordinates: N_0 = 1 - sum xi, N_i = xi_{i-1}; one point (the centroid), exact for the P1 stiffness matrix
    for (int d = 0; d < DIM; d++) for (int k = 0; k < DIM; k++) J[k][d] = X[k + 1][d] - X[0][d];
    const double det = fe_det_inv<DIM>(J, Ji);
    w = fabs(det) / ((DIM == 3) ? 6.0 : 2.0);
    for (int i = 0; i < NC; i++) N[i] = 1.0 / NC;
    for (int d = 0; d < DIM; d++) {
      double s = 0;
      for (int k = 0; k < DIM; k++) { G[k + 1][d] = Ji[d][k]; s += Ji[d][k]; }
      G[0][d] = -s;
    }
  } else {
    // tensor-product element, 2-point Gauss per direction; corner i sits at (bx, by, bz) of the reference cube (UG's corner numbering)
    const double g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);
    double xi[3];
    if constexpr (DIM == 3) { xi[0] = (q & 4) ? g1 : g0; xi[1] = (q & 2) ? g1 : g0; xi[2] = (q & 1) ? g1 : g0; }
    else { xi[0] = (q & 2) ? g1 : g0; xi[1] = (q & 1) ? g1 : g0; xi[2] = g0; }
    double dN[NC][DIM];
    for (int i = 0; i < NC; i++) {
      const int loc[3] = {((i + 1) >> 1) & 1, (i >> 1) & 1, (i >> 2) & 1};
      double f[DIM], df[DIM];
      for (int d = 0; d < DIM; d++) { f[d] = loc[d] ? xi[d] : 1.0 - xi[d]; df[d] = loc[d] ? 1.0 : -1.0; }
      N[i] = 1.0;
      for (int d = 0; d < DIM; d++) N[i] *= f[d];
      for (int d = 0; d < DIM; d++) {
        dN[i][d] = df[d];
        for (int e = 0; e < DIM; e++) if (e != d) dN[i][d] *= f[e];
      }
    }
    for (int k = 0; k < DIM; k++) for (int d = 0; d < DIM; d++) { J[k][d] = 0; for (int i = 0; i < NC; i++) J[k][d] += dN[i][k] * X[i][d]; }
    const double det = fe_det_inv<DIM>(J, Ji);
    w = fabs(det) / ((DIM == 3) ? 8.0 : 4.0);
    for (int i = 0; i < NC; i++) for (int d = 0; d < DIM; d++) { double s = 0; for (int k = 0; k < DIM; k++) s += Ji[d][k] * dN[i][k]; G[i][d] = s; }
  }
}

// local row of corner `ci` of one element: mrow[j][a*BS+b] = sum over the quadrature points (from 0.0, like the zeroed local matrix of
// LocalAssemble) of coefficient * weight * integrand, def[a] likewise
template <int BS, int DIM, int NC>
__device__ __forceinline__ void fe_local_row(const int32_t *__restrict__ er, const double *__restrict__ coord, double kappa, const FeParam &p, int ci,
                                             double (&mrow)[8][BS * BS], double (&def)[BS])
{
  double X[NC][DIM];
  for (int i = 0; i < NC; i++) for (int d = 0; d < DIM; d++) X[i][d] = coord[(size_t)er[i] * DIM + d];
  constexpr int NQ = (NC == DIM + 1) ? 1 : (1 << DIM);
  for (int j = 0; j < NC; j++) for (int k = 0; k < BS * BS; k++) mrow[j][k] = 0.0;
  for (int a = 0; a < BS; a++) def[a] = 0.0;
  for (int q = 0; q < NQ; q++) {
    double w, N[NC], G[NC][DIM];
    fe_qp<DIM, NC>(X, q, w, N, G);
    const double wk = kappa * w;
    const double wn = w * N[ci];
    for (int a = 0; a < BS; a++) def[a] += wn * p.source[a];
    for (int j = 0; j < NC; j++) {
      double dot = 0;
      for (int d = 0; d < DIM; d++) dot += G[ci][d] * G[j][d];
      if constexpr (BS == 1) mrow[j][0] += wk * dot;
      else
        for (int a = 0; a < BS; a++) for (int b = 0; b < BS; b++) {
          const double k = p.lam * G[ci][a] * G[j][b] + p.mu * G[ci][b] * G[j][a] + ((a == b) ? p.mu * dot : 0.0);
          mrow[j][a * BS + b] += wk * k;
        }
    }
  }
}

#define FE_ACC 32
template <int BS, int DIM>
__global__ void __launch_bounds__(128) k_fe_assemble(SellView A, double *aval, const int64_t *__restrict__ iptr, const int64_t *__restrict__ inc,
                                                     const int64_t *__restrict__ eptr, const int32_t *__restrict__ erow, const double *__restrict__ coef,
                                                     const double *__restrict__ coord, const uint32_t *__restrict__ skip, const double *__restrict__ x,
                                                     double *__restrict__ b, FeParam p, int *err)
{
  constexpr int BB = BS * BS;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  const int len = A.rowlen[r];
  const ColIter ci = col_iter(A, r);
  double *ar = aval + slice_off(A, r >> 5) * BB + (r & 31);          // component k of entry t of this row: ar[(t*BB + k)*32]
  // rows of up to FE_ACC entries accumulate in a thread-private array; longer ones (irregular refinement) in place
  const bool local = len <= FE_ACC;
  double acc[FE_ACC * BB];
  int ccol[FE_ACC];
  if (local) {
    for (int t = 0; t < len; t++) { ccol[t] = col_at(ci, t); for (int k = 0; k < BB; k++) acc[t * BB + k] = 0.0; }
  } else {
    for (int t = 0; t < len; t++) for (int k = 0; k < BB; k++) ar[((size_t)t * BB + k) * 32] = 0.0;       // dmatset(A, 0)
  }
  double rhs[BS];
  for (int a = 0; a < BS; a++) rhs[a] = 0.0;                                                               // dset(b, 0)
  for (int64_t q = iptr[r]; q < iptr[r + 1]; q++) {
    const int64_t e = inc[q] >> 3;
    const int c = (int)(inc[q] & 7);
    const int32_t *er = erow + eptr[e];
    const int nc = (int)(eptr[e + 1] - eptr[e]);
    const double kappa = coef ? coef[e] : 1.0;
    double mrow[8][BB], def[BS];
    if (nc == DIM + 1) fe_local_row<BS, DIM, DIM + 1>(er, coord, kappa, p, c, mrow, def);
    else if (nc == (1 << DIM)) fe_local_row<BS, DIM, (1 << DIM)>(er, coord, kappa, p, c, mrow, def);
    else { atomicExch(err, UGGPU_ERROR); continue; }
    for (int a = 0; a < BS; a++) rhs[a] = rhs[a] + def[a];
    for (int j = 0; j < nc; j++) {
      const int w = er[j];
      int t = -1;                                                      // GetMatrix(v, w), disctools.cc:1156
      if (local) { for (int u = 0; u < len; u++) if (ccol[u] == w) { t = u; break; } }
      else { for (int u = 0; u < len; u++) if (col_at(ci, u) == w) { t = u; break; } }
      if (t < 0) { atomicExch(err, UGGPU_DESC_MISMATCH); continue; }   // GetElementVVMPtrs returns -3
      for (int k = 0; k < BB; k++) {
        if (local) acc[t * BB + k] = acc[t * BB + k] + mrow[j][k];
        else ar[((size_t)t * BB + k) * 32] = ar[((size_t)t * BB + k) * 32] + mrow[j][k];
      }
    }
  }
  // AssembleDirichletBoundary: the diagonal block is entry 0 of the row (VSTART)
  const uint32_t sk = skip[r];
  for (int a = 0; a < BS; a++)
    if (sk & (1u << a)) {
      rhs[a] = x[(size_t)r * BS + a];
      for (int t = 0; t < len; t++)
        for (int c = 0; c < BS; c++) {
          const double v = (t == 0 && c == a) ? 1.0 : 0.0;
          if (local) acc[t * BB + a * BS + c] = v; else ar[((size_t)t * BB + a * BS + c) * 32] = v;
        }
    }
  if (local)
    for (int t = 0; t < len; t++) for (int k = 0; k < BB; k++) ar[((size_t)t * BB + k) * 32] = acc[t * BB + k];
  for (int a = 0; a < BS; a++) b[(size_t)r * BS + a] = rhs[a];
}

// (row, element * 8 + corner) of every corner of every element, elements ascending; cnt[row] = incidences
__global__ void k_fe_pairs(int64_t nelem, const int64_t *__restrict__ eptr, const int32_t *__restrict__ erow, int n, int32_t *__restrict__ key,
                           int64_t *__restrict__ val, int *__restrict__ cnt, int *err)
{
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelem) return;
  const int64_t p0 = eptr[e];
  const int nc = (int)(eptr[e + 1] - p0);
  if (nc < 1 || nc > 8) { atomicExch(err, UGGPU_ERROR); return; }
  for (int i = 0; i < nc; i++) {
    const int r = erow[p0 + i];
    if (r < 0 || r >= n) { atomicExch(err, UGGPU_ERROR); key[p0 + i] = 0; val[p0 + i] = e * 8 + i; continue; }
    key[p0 + i] = r; val[p0 + i] = e * 8 + i;
    atomicAdd(&cnt[r], 1);
  }
}

__global__ void k_fe_cnt64(int n, const int *__restrict__ cnt, int64_t *__restrict__ out)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n) out[r] = r < n ? cnt[r] : 0;
}

extern "C" int uggpu_assemble(uggpu_ctx *ctx, int level, int x, int b, int A, const uggpu_fe_cfg *cfg, int64_t nelem, const int64_t *elem_ptr,
                              const int32_t *elem_row, const double *coef, const double *coord, const uint32_t *skip)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!cfg || !elem_ptr || !elem_row || !coord) return uggpu_fail(UGGPU_ERROR, "uggpu_assemble: null argument");
  if (L->partitioned) return uggpu_fail(UGGPU_ERROR, "uggpu_assemble runs on one GPU (level %d is partitioned)", level);
  if (cfg->dim != 2 && cfg->dim != 3) return uggpu_fail(UGGPU_ERROR, "uggpu_assemble: dim %d", cfg->dim);
  if (cfg->problem == UGGPU_FE_POISSON ? L->bs != 1 : (cfg->problem != UGGPU_FE_ELASTICITY || L->bs != cfg->dim))
    return uggpu_fail(UGGPU_DESC_MISMATCH, "uggpu_assemble: problem %d needs %s component(s) per vector, level %d has %d", cfg->problem,
                      cfg->problem == UGGPU_FE_POISSON ? "1" : "dim", level, L->bs);
  SellMat *Am = get_mat(ctx, level, A);
  if (!Am) return UGGPU_DESC_MISMATCH;
  double *xv = get_vec(ctx, level, x), *bv = get_vec(ctx, level, b);
  if (!xv || !bv) return UGGPU_DESC_MISMATCH;
  const int n = L->n, bs = L->bs, dim = cfg->dim;
  if (n == 0) return 0;
  cudaStream_t st = ctx->stream;
  const int64_t ncorner = nelem > 0 ? elem_ptr[nelem] : 0;
  const size_t ne1 = (size_t)nelem + 1, zc = (size_t)(ncorner > 0 ? ncorner : 1);
  int64_t *d_eptr = nullptr, *d_iptr = nullptr, *d_cnt64 = nullptr, *d_val = nullptr, *d_inc = nullptr;
  int32_t *d_erow = nullptr, *d_key = nullptr, *d_key2 = nullptr;
  double *d_coef = nullptr, *d_coord = nullptr;
  int *d_cnt = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0, tb2 = 0;
  int rc = 0;
#define GT(expr) do { if (!rc) rc = (expr); } while (0)
#define GC(expr) do { if (!rc) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); } } while (0)
  GT(dalloc(ctx, &d_eptr, ne1)); GT(dalloc(ctx, &d_erow, zc)); GT(dalloc(ctx, &d_coord, (size_t)n * dim));
  if (coef) GT(dalloc(ctx, &d_coef, (size_t)(nelem > 0 ? nelem : 1)));
  GT(dalloc(ctx, &d_key, zc)); GT(dalloc(ctx, &d_key2, zc)); GT(dalloc(ctx, &d_val, zc)); GT(dalloc(ctx, &d_inc, zc));
  GT(dalloc(ctx, &d_cnt, (size_t)n)); GT(dalloc(ctx, &d_cnt64, (size_t)n + 1)); GT(dalloc(ctx, &d_iptr, (size_t)n + 1));
  if (!rc) {
    GC(cudaMemcpyAsync(d_eptr, elem_ptr, sizeof(int64_t) * ne1, cudaMemcpyHostToDevice, st));
    if (ncorner) GC(cudaMemcpyAsync(d_erow, elem_row, sizeof(int32_t) * (size_t)ncorner, cudaMemcpyHostToDevice, st));
    GC(cudaMemcpyAsync(d_coord, coord, sizeof(double) * (size_t)n * dim, cudaMemcpyHostToDevice, st));
    if (coef && nelem) GC(cudaMemcpyAsync(d_coef, coef, sizeof(double) * (size_t)nelem, cudaMemcpyHostToDevice, st));
    // SetElementDirichletFlags: the VECSKIP words the element loop leaves are the caller's (which components of which vectors are
    // Dirichlet is the application's decision); they become the level's flags like in the reference
    if (skip) GC(cudaMemcpyAsync(L->skip, skip, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
    else GC(cudaMemsetAsync(L->skip, 0, sizeof(uint32_t) * (size_t)n, st));
    int bits = 1;
    while ((1ll << bits) < (long long)n && bits < 31) bits++;
    GC(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt64, d_iptr, n + 1, st));
    GC(cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_key, d_key2, d_val, d_inc, ncorner, 0, bits, st));
    if (tb2 > tmp_bytes) tmp_bytes = tb2;
    GT(dev_alloc(ctx, &tmp, tmp_bytes ? tmp_bytes : 1));
    GC(cudaMemsetAsync(d_cnt, 0, sizeof(int) * (size_t)n, st));
    if (!rc && nelem > 0) { k_fe_pairs<<<(unsigned)((nelem + 255) / 256), 256, 0, st>>>(nelem, d_eptr, d_erow, n, d_key, d_val, d_cnt, ctx->derr); ctx->launches++; }
    if (ncorner > 0) GC(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d_key, d_key2, d_val, d_inc, ncorner, 0, bits, st));   // stable: elements stay ascending per row
    if (!rc) { k_fe_cnt64<<<(n + 256) / 256, 256, 0, st>>>(n, d_cnt, d_cnt64); ctx->launches++; }
    GC(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_cnt64, d_iptr, n + 1, st));
    // the values are about to change: nothing derived from them may survive
    GT(sell_drop_shared_values(ctx, Am));
    GT(sell_free_schedules(ctx, Am));
    if (!rc) {
      FeParam p;
      p.problem = cfg->problem;
      p.lam = cfg->E * cfg->nu / ((1 + cfg->nu) * (1 - 2 * cfg->nu));
      p.mu = cfg->E / (2 * (1 + cfg->nu));
      for (int a = 0; a < UGGPU_MAX_BS; a++) p.source[a] = cfg->source[a];
      const int blocks = (n + 127) / 128;
      const SellView Av = view(*Am);
      ProfScope ps(ctx, UGGPU_K_ASSEMBLE, level, 0.0);
#define FE_LAUNCH(BS_, DIM_) k_fe_assemble<BS_, DIM_><<<blocks, 128, 0, st>>>(Av, Am->val, d_iptr, d_inc, d_eptr, d_erow, d_coef, d_coord, L->skip, xv, bv, p, ctx->derr)
      if (bs == 1 && dim == 2) FE_LAUNCH(1, 2);
      else if (bs == 1) FE_LAUNCH(1, 3);
      else if (bs == 2) FE_LAUNCH(2, 2);
      else FE_LAUNCH(3, 3);
#undef FE_LAUNCH
      ctx->launches++;
      GC(cudaGetLastError());
    }
    bool launched = false;
    if (!rc) {
      launched = true;
      rc = check_device_error(ctx);
      if (rc == UGGPU_DESC_MISMATCH) rc = uggpu_fail(UGGPU_DESC_MISMATCH, "uggpu_assemble: two corners of an element have no matrix entry on level %d (GetElementVVMPtrs -3)", level);
      else if (rc) rc = uggpu_fail(UGGPU_ERROR, "uggpu_assemble: bad element list on level %d (corner rows out of range, or an element that is neither a simplex nor a tensor element)", level);
    }
    // whatever the kernel reported, the values have changed: the diagonal array and the value generation follow them (nothing derived
    // from the old values may be used with the new ones, also after a failed call)
    if (launched) {
      const int rc2 = sell_update_diag(ctx, Am);
      if (!rc) rc = rc2;
      if (!rc) rc = sell_share_values(ctx, Am);
    }
    cudaStreamSynchronize(st);
  }
#undef GT
#undef GC
  if (tmp) dev_free(ctx, tmp, tmp_bytes ? tmp_bytes : 1);
  if (d_eptr) dfree(ctx, d_eptr, ne1);
  if (d_erow) dfree(ctx, d_erow, zc);
  if (d_coord) dfree(ctx, d_coord, (size_t)n * dim);
  if (d_coef) dfree(ctx, d_coef, (size_t)(nelem > 0 ? nelem : 1));
  if (d_key) dfree(ctx, d_key, zc);
  if (d_key2) dfree(ctx, d_key2, zc);
  if (d_val) dfree(ctx, d_val, zc);
  if (d_inc) dfree(ctx, d_inc, zc);
  if (d_cnt) dfree(ctx, d_cnt, (size_t)n);
  if (d_cnt64) dfree(ctx, d_cnt64, (size_t)n + 1);
  if (d_iptr) dfree(ctx, d_iptr, (size_t)n + 1);
  return rc;
}
