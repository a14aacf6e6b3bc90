// spmv.cu -- block-sparse matrix-vector kernels (np/algebra/ugblas.cc:3782-4042 dmatmul, dmatmul_add,
// dmatmul_minus), the diagonal-block solve of l_jac (np/algebra/ugiter.cc:271-335, block.cc:104-142) and the
// fused smoothing step of the multigrid cycle (np/procs/iter.cc:817-842 Smoother + :7811-7816 Lmgc loop body).
//
// One thread per block row on SELL-32 storage: the warp's loads of column indices and values are contiguous
// (128 B / 256 B per instruction), while each thread adds its row's terms in the canonical VSTART->MNEXT
// order with separate multiply and add (-fmad=false), i.e. performs exactly the additions of
// matloop.ct:74-100 / MATMUL_nn_SUCC (ugblas.h:161-282).  Results are therefore bit-identical to the CPU
// reference; tensor cores are not used (0.17-0.25 flop/byte: HBM-bound).
#include "uggpu_internal.h"

#ifndef SPMV_THREADS
#define SPMV_THREADS 128
#endif

template <int BS>
__device__ __forceinline__ void row_product(const SellView &A, int r, const double *__restrict__ y, double (&s)[BS], double (&dg)[BS * BS])
{
  constexpr int BB = BS * BS;
  const int lane = r & 31;
  const int64_t sp = A.slice_ptr[r >> 5];
  const int len = A.rowlen[r];
  const int32_t *__restrict__ cp = A.col + sp + lane;
  const double *__restrict__ vp = A.val + sp * BB + lane;
#pragma unroll
  for (int i = 0; i < BS; i++) s[i] = 0.0;
#pragma unroll
  for (int k = 0; k < BB; k++) dg[k] = 0.0;
#pragma unroll 4
  for (int j = 0; j < len; j++) {
    const int c = __ldg(cp + (size_t)j * 32);
    double m[BB], w[BS];
#pragma unroll
    for (int k = 0; k < BB; k++) m[k] = __ldg(vp + ((size_t)j * BB + k) * 32);
#pragma unroll
    for (int i = 0; i < BS; i++) w[i] = y[(size_t)c * BS + i];
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < BB; k++) dg[k] = m[k];
    }
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double acc = m[i * BS] * w[0];
#pragma unroll
      for (int jj = 1; jj < BS; jj++) acc = acc + m[i * BS + jj] * w[jj];
      s[i] += acc;
    }
  }
}

// SolveSmallBlock (block.cc:104-142), n = 1,2,3.  Returns non-zero for a singular 2x2 block.
template <int BS>
__device__ __forceinline__ int solve_small_block(const double (&mat)[BS * BS], const double (&rhs)[BS], double (&sol)[BS])
{
  if (BS == 1) { sol[0] = rhs[0] / mat[0]; return 0; }
  if (BS == 2) {
    double det = mat[0] * mat[3 % (BS * BS)] - mat[1 % (BS * BS)] * mat[2 % (BS * BS)];
    if (det == 0.0) return 1;
    det = 1.0 / det;
    sol[0] = (rhs[0] * mat[3 % (BS * BS)] - rhs[1 % BS] * mat[1 % (BS * BS)]) * det;
    sol[1 % BS] = (rhs[1 % BS] * mat[0] - rhs[0] * mat[2 % (BS * BS)]) * det;
    return 0;
  }
  // n == 3 (indices wrapped with % only to keep the BS<3 instantiations well-formed)
  constexpr int BB = BS * BS;
  double M3div0 = mat[3 % BB] / mat[0];
  double M6div0 = mat[6 % BB] / mat[0];
  double aux = (mat[7 % BB] - M6div0 * mat[1 % BB]) / (mat[4 % BB] - M3div0 * mat[1 % BB]);
  sol[2 % BS] = (rhs[2 % BS] - M6div0 * rhs[0] - aux * (rhs[1 % BS] - M3div0 * rhs[0]))
                / (mat[8 % BB] - M6div0 * mat[2 % BB] - aux * (mat[5 % BB] - M3div0 * mat[2 % BB]));
  sol[1 % BS] = (rhs[1 % BS] - mat[3 % BB] / mat[0] * rhs[0] - (mat[5 % BB] - M3div0 * mat[2 % BB]) * sol[2 % BS])
                / (mat[4 % BB] - M3div0 * mat[1 % BB]);
  sol[0] = (rhs[0] - mat[1 % BB] * sol[1 % BS] - mat[2 % BB] * sol[2 % BS]) / mat[0];
  return 0;
}

// ---- dmatmul family ---------------------------------------------------------------------------------------------
template <int BS, int OP>
__global__ void __launch_bounds__(SPMV_THREADS) k_dmatmul_k(SellView A, uint8_t bit, const uint8_t *__restrict__ ctl, double *__restrict__ x, const double *__restrict__ y)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  if (bit && !(ctl[r] & bit)) return;
  double s[BS], dg[BS * BS];
  row_product<BS>(A, r, y, s, dg);
#pragma unroll
  for (int i = 0; i < BS; i++) {
    size_t k = (size_t)r * BS + i;
    if (OP == 0) x[k] = (BS == 1) ? s[i] : 0.0 + s[i];     // matmode.ct:76 T_CLEAR_X then +=
    else if (OP == 1) x[k] = x[k] + s[i];
    else x[k] = x[k] - s[i];
  }
}

template <int BS>
static int launch_dmatmul(uggpu_ctx *ctx, Level *L, const SellMat *A, int op, int rowmode, double *x, const double *y)
{
  if (L->n == 0) return 0;
  uint8_t bit = rowmode == 0 ? 0 : (rowmode == 1 ? UGGPU_CTL_NEW_DEFECT : UGGPU_CTL_FINE_GRID_DOF);
  int blocks = (L->n + SPMV_THREADS - 1) / SPMV_THREADS;
  SellView v = view(*A);
  const double nb = 8.0 * BS * L->n;
  ProfScope ps(ctx, UGGPU_K_DMATMUL, (int)(L - ctx->lev), (double)A->nnz * (8.0 * BS * BS + 4.0) + 4.0 * (L->n + 1.0) + (op == 0 ? 2.0 : 3.0) * nb);
  if (op == 0) k_dmatmul_k<BS, 0><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y);
  else if (op == 1) k_dmatmul_k<BS, 1><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y);
  else k_dmatmul_k<BS, 2><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y);
  KCHECK(ctx);
  return 0;
}

int k_dmatmul(uggpu_ctx *ctx, int level, int op, int rowmode, int x, int M, int y)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  SellMat *A = get_mat(ctx, level, M);
  double *xp = get_vec(ctx, level, x);
  const double *yp = get_vec(ctx, level, y);
  if (!A || !xp || !yp) return UGGPU_DESC_MISMATCH;
  if (xp == yp) return uggpu_fail(UGGPU_DESC_MISMATCH, "dmatmul: result and operand are the same vector");
  UG_TRY(halo_exchange(ctx, level, const_cast<double *>(yp)));   // ghost columns of the operand (no-op on one GPU)
  switch (L->bs) {
    case 1: return launch_dmatmul<1>(ctx, L, A, op, rowmode, xp, yp);
    case 2: return launch_dmatmul<2>(ctx, L, A, op, rowmode, xp, yp);
    default: return launch_dmatmul<3>(ctx, L, A, op, rowmode, xp, yp);
  }
}

static int matmul_loop(uggpu_ctx *ctx, int fl, int tl, int mode, int op, int x, int M, int y)
{
  std::vector<LoopItem> items;
  UG_TRY(surface_loop(ctx, fl, tl, mode, items));
  for (auto &it : items) UG_TRY(k_dmatmul(ctx, it.level, op, it.rowmode, x, M, y));
  return 0;
}

extern "C" int uggpu_dmatmul(uggpu_ctx *c, int fl, int tl, int mode, int x, int M, int y) { return matmul_loop(c, fl, tl, mode, 0, x, M, y); }
extern "C" int uggpu_dmatmul_add(uggpu_ctx *c, int fl, int tl, int mode, int x, int M, int y) { return matmul_loop(c, fl, tl, mode, 1, x, M, y); }
extern "C" int uggpu_dmatmul_minus(uggpu_ctx *c, int fl, int tl, int mode, int x, int M, int y) { return matmul_loop(c, fl, tl, mode, 2, x, M, y); }

// ---- l_jac ----------------------------------------------------------------------------------------------------------
// v = damp (.) Diag(A)^-1 d, v = 0 where VCLASS < ACTIVE_CLASS (ugiter.cc:300).  The diagonal block is entry 0
// of the row, i.e. the first 32-wide column of the slice: a coalesced read.
template <int BS>
__global__ void __launch_bounds__(SPMV_THREADS) k_jac_k(SellView A, const uint8_t *__restrict__ vclass, double *__restrict__ v, const double *__restrict__ d, Damp damp, int *err)
{
  constexpr int BB = BS * BS;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  double sol[BS];
  if (vclass[r] < 3) {
#pragma unroll
    for (int i = 0; i < BS; i++) sol[i] = 0.0;
  } else {
    const int64_t sp = A.slice_ptr[r >> 5];
    const double *__restrict__ vp = A.val + sp * BB + (r & 31);
    double m[BB], rhs[BS];
#pragma unroll
    for (int k = 0; k < BB; k++) m[k] = vp[(size_t)k * 32];
#pragma unroll
    for (int i = 0; i < BS; i++) rhs[i] = d[(size_t)r * BS + i];
    if (solve_small_block<BS>(m, rhs, sol)) { atomicExch(err, UGGPU_SMALL_DIAG); return; }
  }
#pragma unroll
  for (int i = 0; i < BS; i++) v[(size_t)r * BS + i] = sol[i] * damp.a[i];
}

int k_jac(uggpu_ctx *ctx, int level, int A, double *v, const double *d, Damp damp)
{
  Level *L = get_level(ctx, level);
  SellMat *M = get_mat(ctx, level, A);
  if (!L || !M) return UGGPU_DESC_MISMATCH;
  if (L->n == 0) return 0;
  int blocks = (L->n + SPMV_THREADS - 1) / SPMV_THREADS;
  SellView vw = view(*M);
  ProfScope ps(ctx, UGGPU_K_JAC, level, (double)L->n * (8.0 * L->bs * L->bs + 16.0 * L->bs));
  switch (L->bs) {
    case 1: k_jac_k<1><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr); break;
    case 2: k_jac_k<2><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr); break;
    default: k_jac_k<3><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr); break;
  }
  KCHECK(ctx);
  return 0;
}

extern "C" int uggpu_l_jac(uggpu_ctx *ctx, int level, int v, int M, int d)
{
  double *vp = get_vec(ctx, level, v);
  const double *dp = get_vec(ctx, level, d);
  if (!vp || !dp) return UGGPU_DESC_MISMATCH;
  Damp one = mkdamp(nullptr, 0);        // sol * 1.0 is exact
  UG_TRY(k_jac(ctx, level, M, vp, dp, one));
  return check_device_error(ctx);
}

// Smoother() (iter.cc:817-842) with Step = JacobiStep (:911), one kernel per reference call:
// l_jac ; dscalx(damp) ; dmatmul_minus
extern "C" int uggpu_jac_smooth(uggpu_ctx *ctx, int level, int x, int b, int A, const double *damp)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  double *xp = get_vec(ctx, level, x);
  double *bp = get_vec(ctx, level, b);
  if (!xp || !bp) return UGGPU_DESC_MISMATCH;
  UG_TRY(k_jac(ctx, level, A, xp, bp, mkdamp(nullptr, 0)));
  UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, xp, nullptr, mkdamp(damp, L->bs)));
  UG_TRY(k_dmatmul(ctx, level, 2, 0, b, A, x));
  return 0;
}

// ---- fused smoothing step ----------------------------------------------------------------------------------------------
// For row r, with tin = the damped Jacobi correction of this step (already computed for ALL rows):
//     b[r]  -= (A tin)[r]                    dmatmul_minus  iter.cc:838
//     c[r]  += tin[r]   (or 0 + tin[r])      dadd           iter.cc:7814
//     tout[r] = damp * Diag(A)^-1 b[r]       l_jac + dscalx of the NEXT step (iter.cc:911,836), class-masked
//     x[r]  += c[r]                          LSUpdate       ls.cc:869      (last step of the top level)
//     partial sums of b[r]^2 over NEW_DEFECT rows          LinearResiduum ls.cc:577 (ditto)
// Every quantity is produced by the same arithmetic operations on the same operands as in the one-kernel-
// per-call path, so fused and unfused results are bit-identical.  tout must not alias tin (other rows gather tin).
template <int BS, int FLAGS>
__global__ void __launch_bounds__(SPMV_THREADS) k_smooth_k(SellView A, const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl,
                                                           const double *__restrict__ tin, double *__restrict__ b, double *__restrict__ c,
                                                           double *__restrict__ tout, Damp damp, double *__restrict__ x, double *__restrict__ partials, int *err)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = r < A.n;
  double nrm[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) nrm[i] = 0.0;
  if (active) {
    double s[BS], dg[BS * BS], bn[BS];
    row_product<BS>(A, r, tin, s, dg);
#pragma unroll
    for (int i = 0; i < BS; i++) {
      size_t k = (size_t)r * BS + i;
      bn[i] = b[k] - s[i];
      b[k] = bn[i];
    }
    if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
#pragma unroll
      for (int i = 0; i < BS; i++) {
        size_t k = (size_t)r * BS + i;
        double cn;
        if (FLAGS & SF_CADD) cn = c[k] + tin[k];
        else if (FLAGS & SF_CSET) cn = 0.0 + tin[k];
        else cn = c[k];
        if (FLAGS & (SF_CADD | SF_CSET)) c[k] = cn;
        if (FLAGS & SF_XADD) x[k] = x[k] + cn;
      }
    }
    if (FLAGS & SF_TOUT) {
      double sol[BS];
      if (vclass[r] < 3) {
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      } else if (solve_small_block<BS>(dg, bn, sol)) {
        atomicExch(err, UGGPU_SMALL_DIAG);
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      }
#pragma unroll
      for (int i = 0; i < BS; i++) tout[(size_t)r * BS + i] = sol[i] * damp.a[i];
    }
    if (FLAGS & SF_NORM) {
      if (ctl[r] & UGGPU_CTL_NEW_DEFECT) {
#pragma unroll
        for (int i = 0; i < BS; i++) nrm[i] = bn[i] * bn[i];
      }
    }
  }
  if (FLAGS & SF_NORM) {
    __shared__ double sm[SPMV_THREADS / 32][UGGPU_MAX_BS];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double v = nrm[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) sm[w][i] = v;
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int i = 0; i < BS; i++) {
        double v = lane < SPMV_THREADS / 32 ? sm[lane][i] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) partials[(size_t)blockIdx.x * BS + i] = v;
      }
    }
  }
}

template <int BS, int FLAGS>
static int launch_smooth2(uggpu_ctx *ctx, Level *L, const SellMat *A, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot)
{
  int blocks = (L->n + SPMV_THREADS - 1) / SPMV_THREADS;
  if (FLAGS & SF_NORM) UG_TRY(ensure_partials(ctx, (size_t)blocks * BS));
  const double nb = 8.0 * BS * L->n;
  // algorithmic bytes: entries, row lengths, gathered operand once, b read+write, c, tout, x
  ProfScope ps(ctx, UGGPU_K_SMOOTH, (int)(L - ctx->lev), (double)A->nnz * (8.0 * BS * BS + 4.0) + 4.0 * (L->n + 1.0) + 3.0 * nb
               + ((FLAGS & SF_CADD) ? 2.0 * nb : 0.0) + ((FLAGS & SF_CSET) ? nb : 0.0) + ((FLAGS & SF_TOUT) ? nb : 0.0) + ((FLAGS & SF_XADD) ? 2.0 * nb : 0.0));
  k_smooth_k<BS, FLAGS><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr);
  KCHECK(ctx);
  if (FLAGS & SF_NORM) UG_TRY(reduce_partials_final(ctx, BS, (size_t)blocks, norm_slot, (int)(L - ctx->lev)));
  return 0;
}

template <int BS>
static int launch_smooth(uggpu_ctx *ctx, Level *L, const SellMat *A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot)
{
#define SM_CASE(F) case F: return launch_smooth2<BS, F>(ctx, L, A, tin, b, c, tout, damp, x, norm_slot)
  switch (flags) {
    SM_CASE(0);
    SM_CASE(SF_CADD);
    SM_CASE(SF_CSET);
    SM_CASE(SF_TOUT);
    SM_CASE(SF_CADD | SF_TOUT);
    SM_CASE(SF_CSET | SF_TOUT);
    SM_CASE(SF_CADD | SF_XADD | SF_NORM);
    SM_CASE(SF_CSET | SF_XADD | SF_NORM);
    SM_CASE(SF_CADD | SF_XADD);
    SM_CASE(SF_CSET | SF_XADD);
    SM_CASE(SF_CADD | SF_NORM);
    SM_CASE(SF_CSET | SF_NORM);
    SM_CASE(SF_NORM);
  }
#undef SM_CASE
  return uggpu_fail(UGGPU_ERROR, "smooth step: unsupported flag combination %d", flags);
}

int k_smooth_step(uggpu_ctx *ctx, int level, int A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot)
{
  Level *L = get_level(ctx, level);
  SellMat *M = get_mat(ctx, level, A);
  if (!L || !M) return UGGPU_DESC_MISMATCH;
  if (L->n == 0) return 0;
  if ((flags & SF_TOUT) && tout == tin) return uggpu_fail(UGGPU_ERROR, "smooth step: tout aliases tin");
  UG_TRY(halo_exchange(ctx, level, const_cast<double *>(tin)));  // ghost columns of the correction (no-op on one GPU)
  switch (L->bs) {
    case 1: return launch_smooth<1>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot);
    case 2: return launch_smooth<2>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot);
    default: return launch_smooth<3>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot);
  }
}
