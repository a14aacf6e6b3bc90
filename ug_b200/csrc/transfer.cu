// transfer.cu -- standard (geometric) grid transfer of np/algebra/transgrid.cc on precomputed stencils.
//
// The reference recomputes the P1/Q1 weights from the element geometry on every call
// (StandardRestrictNodeVector transgrid.cc:117-215, StandardIntCorNodeVector :235-336).  PreProcess
// flattens them once into P (rows = fine vectors, corner order, zero weights dropped as :304 does) and
// R (rows = coarse vectors, entries in fine NODE list order = the order in which the reference's scatter
// loop :150-189 adds into that coarse vector).  Both are stored SELL-32 (scalar weight per entry), and one
// thread per destination row adds the terms in stored order: gather formulation, no atomics, and the same
// additions in the same order as the reference -> bit-identical results.
#include "uggpu_internal.h"

#define TR_THREADS 256

// StandardRestrict (transgrid.cc:462 -> :117): to[coarse] (zeroed where VNCLASS >= NEWDEF_CLASS) += sum w * (damp*from[fine]),
// suppressed per component by the coarse VECSKIP bits (:161-165,:180-186).
// Optionally fused (FUSE): the first Jacobi correction of the coarse level, tout = sdamp * Diag(Ac)^-1 to (class-masked,
// ugiter.cc:271 + iter.cc:836) and c = 0 (dset, iter.cc:7873).
template <int BS, bool FUSE>
__global__ void __launch_bounds__(TR_THREADS) k_restrict_k(SellView R, const uint8_t *__restrict__ vnclass_c, const uint32_t *__restrict__ skip_c,
                                                           double *__restrict__ to, const double *__restrict__ from, Damp damp,
                                                           SellView Ac, const uint8_t *__restrict__ vclass_c, double *__restrict__ tout, double *__restrict__ czero,
                                                           Damp sdamp, int *err, Prefetch pf)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R.n) return;
  const PfState pfs = pf_begin(R, r, pf);
  double tr[BS];
  const bool zero = vnclass_c[r] >= 2;
#pragma unroll
  for (int i = 0; i < BS; i++) tr[i] = zero ? 0.0 : to[(size_t)r * BS + i];
  const uint32_t skip = skip_c[r];
  const int lane = r & 31;
  const int64_t sp = R.slice_ptr[r >> 5];
  const int len = R.rowlen[r];
  const ColIter ci = col_iter(R, r);
  const double *__restrict__ wp = R.val + sp + lane;
#pragma unroll 4
  for (int j = 0; j < len; j++) {
    const int f = col_at(ci, j);
    const double w = __ldg(wp + (size_t)j * 32);
#pragma unroll
    for (int i = 0; i < BS; i++) {
      if (!(skip & (1u << i))) {
        double s = damp.a[i] * from[(size_t)f * BS + i];
        tr[i] = tr[i] + w * s;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < BS; i++) to[(size_t)r * BS + i] = tr[i];
  pf_end<1>(R, pfs, pf);
  if ((pf.mode & 4) && pfs.sp >= 0) { pf_rows<1>(vnclass_c, pfs, pf); pf_rows<4>(skip_c, pfs, pf); if (FUSE) pf_rows<1>(vclass_c, pfs, pf); }
  if (FUSE) {
    constexpr int BB = BS * BS;
    double sol[BS];
    if (vclass_c[r] < 3) {
#pragma unroll
      for (int i = 0; i < BS; i++) sol[i] = 0.0;
    } else {
      const double *__restrict__ vp = Ac.diag + ((size_t)(r >> 5) * BB) * 32 + lane;
      double m[BB];
#pragma unroll
      for (int k = 0; k < BB; k++) m[k] = vp[(size_t)k * 32];
      if (BS == 1) sol[0] = tr[0] / m[0];
      else {
        // same closed forms as solve_small_block (spmv.cu); duplicated here to keep the kernels self-contained
        if (BS == 2) {
          double det = m[0] * m[3 % BB] - m[1 % BB] * m[2 % BB];
          if (det == 0.0) { atomicExch(err, UGGPU_SMALL_DIAG); det = 1.0; }
          det = 1.0 / det;
          sol[0] = (tr[0] * m[3 % BB] - tr[1 % BS] * m[1 % BB]) * det;
          sol[1 % BS] = (tr[1 % BS] * m[0] - tr[0] * m[2 % BB]) * det;
        } else {
          double M3div0 = m[3 % BB] / m[0];
          double M6div0 = m[6 % BB] / m[0];
          double aux = (m[7 % BB] - M6div0 * m[1 % BB]) / (m[4 % BB] - M3div0 * m[1 % BB]);
          sol[2 % BS] = (tr[2 % BS] - M6div0 * tr[0] - aux * (tr[1 % BS] - M3div0 * tr[0]))
                        / (m[8 % BB] - M6div0 * m[2 % BB] - aux * (m[5 % BB] - M3div0 * m[2 % BB]));
          sol[1 % BS] = (tr[1 % BS] - m[3 % BB] / m[0] * tr[0] - (m[5 % BB] - M3div0 * m[2 % BB]) * sol[2 % BS])
                        / (m[4 % BB] - M3div0 * m[1 % BB]);
          sol[0] = (tr[0] - m[1 % BB] * sol[1 % BS] - m[2 % BB] * sol[2 % BS]) / m[0];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < BS; i++) {
      tout[(size_t)r * BS + i] = sol[i] * sdamp.a[i];
      czero[(size_t)r * BS + i] = 0.0;
    }
  }
}

// StandardInterpolateCorrection (transgrid.cc:529 -> :235): to[fine] = sum (w*damp) * from[coarse], components with the fine
// VECSKIP bit set stay 0 (:272-285).
template <int BS>
__global__ void __launch_bounds__(TR_THREADS) k_interpolate_k(SellView P, const uint32_t *__restrict__ skip_f, double *__restrict__ to,
                                                              const double *__restrict__ from, Damp damp, Prefetch pf)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P.n) return;
  const PfState pfs = pf_begin(P, r, pf);
  double tr[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) tr[i] = 0.0;
  const uint32_t skip = skip_f[r];
  const int lane = r & 31;
  const int64_t sp = P.slice_ptr[r >> 5];
  const int len = P.rowlen[r];
  const ColIter ci = col_iter(P, r);
  const double *__restrict__ wp = P.val + sp + lane;
#pragma unroll 2
  for (int j = 0; j < len; j++) {
    const int c = col_at(ci, j);
    const double w = __ldg(wp + (size_t)j * 32);
#pragma unroll
    for (int i = 0; i < BS; i++)
      if (!(skip & (1u << i))) tr[i] = tr[i] + (w * damp.a[i]) * from[(size_t)c * BS + i];
  }
#pragma unroll
  for (int i = 0; i < BS; i++) to[(size_t)r * BS + i] = tr[i];
  pf_end<1>(P, pfs, pf);
  if ((pf.mode & 4) && pfs.sp >= 0) pf_rows<4>(skip_f, pfs, pf);
}

int k_restrict(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp, bool fuse, int A, double *tout, double *czero, Damp sdamp)
{
  Level *F = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!F || !C) return UGGPU_NO_COARSER_GRID;
  if (!F->R.valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "no transfer stencils on level %d (uggpu_transfer_set)", level);
  if (C->n == 0) return 0;
  // partitioned fine level: the coarse rows this rank owns gather from fine ghost rows too
  UG_TRY(halo_exchange(ctx, level, const_cast<double *>(from)));
  const bool gather = ctx->comm && F->partitioned && !C->partitioned;   // first completely held (replicated) level
  if (gather && fuse) return uggpu_fail(UGGPU_ERROR, "restrict: fused Jacobi start not possible across the gather level");
  int blocks = (C->n + TR_THREADS - 1) / TR_THREADS;
  SellView Rv = view(F->R);
  SellView Av = Rv;
  if (fuse) {
    SellMat *M = get_mat(ctx, level - 1, A);
    if (!M) return UGGPU_DESC_MISMATCH;
    Av = view(*M);
  }
 const double nbf = 8.0 * F->bs * F->n, nbc = 8.0 * F->bs * C->n;
  ProfScope ps(ctx, UGGPU_K_RESTRICT, level, F->R.entry_bytes() + 4.0 * (C->n + 1.0) + nbf + nbc + (fuse ? (double)C->n * 8.0 * F->bs * F->bs + 2.0 * nbc : 0.0));
  Prefetch pf = make_prefetch(ctx, &F->R, F->bs);
  pf.val_lines = (F->R.maxlen * 256 + 127) / 128;     // scalar weights whatever the block size
#define RS(BSV)                                                                                                                        \
  if (fuse) k_restrict_k<BSV, true><<<blocks, TR_THREADS, 0, ctx->stream>>>(Rv, C->vnclass, C->skip, to, from, damp, Av, C->vclass, tout, czero, sdamp, ctx->derr, pf); \
  else k_restrict_k<BSV, false><<<blocks, TR_THREADS, 0, ctx->stream>>>(Rv, C->vnclass, C->skip, to, from, damp, Av, C->vclass, tout, czero, sdamp, ctx->derr, pf)
  switch (F->bs) {
    case 1: RS(1); break;
    case 2: RS(2); break;
    default: RS(3); break;
  }
#undef RS
  KCHECK(ctx);
  // every rank filled only the rows of the coarse nodes it would own (the others are 0): summing the disjoint parts
  // is the gather of the coarse defect onto every rank (agglomeration)
  if (gather) UG_TRY(allreduce_sum(ctx, to, (size_t)C->n * C->bs));
  return 0;
}

int k_interpolate(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp)
{
  Level *F = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!F || !C) return UGGPU_NO_COARSER_GRID;
  if (!F->P.valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "no transfer stencils on level %d (uggpu_transfer_set)", level);
  if (F->n == 0) return 0;
  UG_TRY(halo_exchange(ctx, level - 1, const_cast<double *>(from)));   // coarse ghost values (no-op if the coarse level is replicated)
  int blocks = (F->n + TR_THREADS - 1) / TR_THREADS;
  ProfScope ps(ctx, UGGPU_K_INTERPOLATE, level, F->P.entry_bytes() + 4.0 * (F->n + 1.0) + 8.0 * F->bs * ((double)F->n + C->n));
  const Prefetch pf = make_prefetch(ctx, &F->P, F->bs);
  switch (F->bs) {
    case 1: k_interpolate_k<1><<<blocks, TR_THREADS, 0, ctx->stream>>>(view(F->P), F->skip, to, from, damp, pf); break;
    case 2: k_interpolate_k<2><<<blocks, TR_THREADS, 0, ctx->stream>>>(view(F->P), F->skip, to, from, damp, pf); break;
    default: k_interpolate_k<3><<<blocks, TR_THREADS, 0, ctx->stream>>>(view(F->P), F->skip, to, from, damp, pf); break;
  }
  KCHECK(ctx);
  return 0;
}

extern "C" int uggpu_restrict(uggpu_ctx *ctx, int level, int to, int from, const double *damp)
{
  Level *F = get_level(ctx, level);
  if (!F) return UGGPU_ERROR;
  double *tp = get_vec(ctx, level - 1, to);
  const double *fp = get_vec(ctx, level, from);
  if (!tp || !fp) return UGGPU_DESC_MISMATCH;
  return k_restrict(ctx, level, tp, fp, mkdamp(damp, F->bs), false, -1, nullptr, nullptr, mkdamp(nullptr, 0));
}

extern "C" int uggpu_interpolate_correction(uggpu_ctx *ctx, int level, int to, int from, const double *damp)
{
  Level *F = get_level(ctx, level);
  if (!F) return UGGPU_ERROR;
  double *tp = get_vec(ctx, level, to);
  const double *fp = get_vec(ctx, level - 1, from);
  if (!tp || !fp) return UGGPU_DESC_MISMATCH;
  return k_interpolate(ctx, level, tp, fp, mkdamp(damp, F->bs));
}
