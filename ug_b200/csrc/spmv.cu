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

#include <cstdlib>
#include <vector>

#ifndef SPMV_THREADS
#define SPMV_THREADS 128
#endif
#ifndef SPMV_UNROLL_U
#define SPMV_UNROLL_U 4      // unroll of the uniform-slice loop (gather addresses are load-independent there)
#endif
#ifndef SPMV_MINBLOCKS
#define SPMV_MINBLOCKS (2048 / SPMV_THREADS)   // scalar rows: resident blocks per SM the smoothing kernel is compiled for (register bound)
#endif
#ifndef SPMV_FULL
#define SPMV_FULL 1          // 1: slices with shared values whose rows all have the full width run a predicate-free loop
#endif
#ifndef SPMV_EARLY
#define SPMV_EARLY 0         // 1: the row's own b, c, tin entries are requested before the row product (scalar rows)
#endif
#define SPMV_PRAGMA_(x) _Pragma(#x)
#define SPMV_PRAGMA(x) SPMV_PRAGMA_(x)

// One row times y.  UNIFORM selects the column-index form of the slice (uggpu_internal.h): explicit words, one per entry,
// or one distance per slice column shared by the 32 rows.  In the uniform form lane j holds distance j in a register (one
// coalesced load per slice) and the loop broadcasts it with a shuffle, so the gather addresses do not depend on a load.
// VSHARED (uniform slices only): the slice's values come from a shared table, one block per slice column (sell.cu
// sell_share_values) -- a warp-uniform, L1-resident load instead of 256 bytes of HBM stream per component and column.
// CG: the gathers bypass L1 (ld.global.cg) -- slices of a partitioned level that read GHOST columns (multi-GPU, HaloK): the ghost rows of
// y are written by other GPUs while the kernel runs; they are complete once the warp has passed halo_wait, but a line that holds the last
// owned rows AND the first ghost rows may already sit in this SM's L1 from another warp's gathers.  L2 is where peer writes land.
template <bool CG> __device__ __forceinline__ double gather_ld(const double *p, bool ghost) { return (CG && ghost) ? __ldcg(p) : *p; }

template <int BS, bool UNIFORM, bool VSHARED = false, bool CG = false>
__device__ __forceinline__ void row_product_t(const SellView &A, int r, int len, int64_t cpo, const double *__restrict__ y, double (&s)[BS], double (&dg)[BS * BS])
{
  constexpr int BB = BS * BS;
  const int lane = r & 31;
  const int64_t sp = slice_off(A, r >> 5);
  const double *__restrict__ vp = A.val + sp * BB + lane;
#pragma unroll
  for (int i = 0; i < BS; i++) s[i] = 0.0;
#pragma unroll
  for (int k = 0; k < BB; k++) dg[k] = 0.0;
  if (UNIFORM) {
    // all 32 lanes of the warp are here (rows that do not take part have len = 0)
    const int w = slice_width(A, r >> 5, sp);
    const int32_t *__restrict__ dp = A.col + UG_COLTAB(cpo);
    const double *__restrict__ tp = VSHARED ? A.vt + UG_VALTAB(cpo) : nullptr;
#if SPMV_FULL
    // all 32 rows have the slice's full width (every slice of interior rows): no per-lane predicate inside the loop, so the
    // compiler is free to batch the loop's loads
    if (VSHARED && w <= 32 && __all_sync(0xffffffffu, len == w)) {
      const int dreg = lane < w ? __ldg(dp + lane) : 0;
SPMV_PRAGMA(unroll SPMV_UNROLL_U)
      for (int j = 0; j < w; j++) {
        const int c = r + __shfl_sync(0xffffffffu, dreg, j);
        double m[BB], wv[BS];
#pragma unroll
        for (int k = 0; k < BB; k++) m[k] = __ldg(tp + (size_t)j * BB + k);
#pragma unroll
        for (int i = 0; i < BS; i++) wv[i] = gather_ld<CG>(y + (size_t)c * BS + i, c >= A.n);
        if (j == 0) {
#pragma unroll
          for (int k = 0; k < BB; k++) dg[k] = m[k];
        }
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = m[i * BS] * wv[0];
#pragma unroll
          for (int q = 1; q < BS; q++) acc = acc + m[i * BS + q] * wv[q];
          s[i] += acc;
        }
      }
      return;
    }
#endif
    for (int j0 = 0; j0 < w; j0 += 32) {
      const int dreg = (j0 + lane < w) ? __ldg(dp + j0 + lane) : 0;
      const int jn = min(32, w - j0);
SPMV_PRAGMA(unroll SPMV_UNROLL_U)
      for (int jj = 0; jj < jn; jj++) {
        const int j = j0 + jj;
        const int c = r + __shfl_sync(0xffffffffu, dreg, jj);
        if (j < len) {
          double m[BB], wv[BS];
#pragma unroll
          for (int k = 0; k < BB; k++) m[k] = VSHARED ? __ldg(tp + (size_t)j * BB + k) : __ldg(vp + ((size_t)j * BB + k) * 32);
#pragma unroll
          for (int i = 0; i < BS; i++) wv[i] = gather_ld<CG>(y + (size_t)c * BS + i, c >= A.n);
          if (j == 0) {
#pragma unroll
            for (int k = 0; k < BB; k++) dg[k] = m[k];
          }
#pragma unroll
          for (int i = 0; i < BS; i++) {
            double acc = m[i * BS] * wv[0];
#pragma unroll
            for (int q = 1; q < BS; q++) acc = acc + m[i * BS + q] * wv[q];
            s[i] += acc;
          }
        }
      }
    }
  } else {
    const int32_t *__restrict__ cp = A.col + cpo + lane;
#pragma unroll 4
    for (int j = 0; j < len; j++) {
      const int c = __ldg(cp + (size_t)j * 32);
      double m[BB], w[BS];
#pragma unroll
      for (int k = 0; k < BB; k++) m[k] = __ldg(vp + ((size_t)j * BB + k) * 32);
#pragma unroll
      for (int i = 0; i < BS; i++) w[i] = gather_ld<CG>(y + (size_t)c * BS + i, c >= A.n);
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
}

// Must be entered by whole warps whose slice exists ((r & ~31) < A.n): the uniform form shuffles across the slice.
// Rows with live == false (beyond n, or masked out by the caller) contribute nothing and get s = 0.
// the two direct-indexed loads every row starts with (code word of the slice, row length)
__device__ __forceinline__ void row_head(const SellView &A, int r, bool live, int64_t &cpo, int &len)
{
  cpo = (A.fixed_w && A.col_ptr == A.slice_ptr) ? (int64_t)(r >> 5) * 32 * A.fixed_w : A.col_ptr[r >> 5];
  len = live ? (int)A.rowlen[r] : 0;
}
template <int BS, bool CG = false>
__device__ __forceinline__ void row_product_h(const SellView &A, int r, int64_t cpo, int len, const double *__restrict__ y, double (&s)[BS], double (&dg)[BS * BS])
{
  if (cpo < 0) {
    if (A.vt && UG_VALTAB(cpo) >= 0) row_product_t<BS, true, true, CG>(A, r, len, cpo, y, s, dg);
    else row_product_t<BS, true, false, CG>(A, r, len, cpo, y, s, dg);
  } else row_product_t<BS, false, false, CG>(A, r, len, cpo, y, s, dg);
}
template <int BS, bool CG = false>
__device__ __forceinline__ void row_product(const SellView &A, int r, bool live, const double *__restrict__ y, double (&s)[BS], double (&dg)[BS * BS])
{
  const int64_t cpo = (A.fixed_w && A.col_ptr == A.slice_ptr) ? (int64_t)(r >> 5) * 32 * A.fixed_w : A.col_ptr[r >> 5];
  const int len = live ? (int)A.rowlen[r] : 0;
  if (cpo < 0) {
    if (A.vt && UG_VALTAB(cpo) >= 0) row_product_t<BS, true, true, CG>(A, r, len, cpo, y, s, dg);
    else row_product_t<BS, true, false, CG>(A, r, len, cpo, y, s, dg);
  } else row_product_t<BS, false, false, CG>(A, r, len, cpo, y, s, dg);
}

// ---- dmatmul family ---------------------------------------------------------------------------------------------
template <int BS, int OP>
__global__ void __launch_bounds__(SPMV_THREADS) k_dmatmul_k(SellView A, uint8_t bit, const uint8_t *__restrict__ ctl, double *__restrict__ x, const double *__restrict__ y,
                                                            Prefetch pf)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if ((r & ~31) >= A.n) return;
  const PfState pfs = pf_begin(A, r, pf);
  const bool live = r < A.n && (!bit || (ctl[r] & bit));
  double s[BS], dg[BS * BS];
  row_product<BS>(A, r, live, y, s, dg);
  pf_end<BS * BS>(A, pfs, pf);
  if ((pf.mode & 4) && pfs.sp >= 0 && OP != 0) pf_vec<BS>(x, pfs, pf);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < BS; i++) {
    size_t k = (size_t)r * BS + i;
    if (OP == 0) x[k] = (BS == 1) ? s[i] : 0.0 + s[i];     // matmode.ct:76 T_CLEAR_X then +=
    else if (OP == 1) x[k] = x[k] + s[i];
    else x[k] = x[k] - s[i];
  }
}

// Stencil variant (scalar rows), the dmatmul counterpart of k_smooth_sten below: on a matrix whose slices mostly carry ONE
// (distance, value) table pair the tables arrive as a kernel parameter (constant bank), W is a template parameter, and a slice of
// full-width rows with that code word runs the unrolled, predicate-free loop; every other slice takes row_product.  Same sums in
// the same order as k_dmatmul_k.
template <int OP, int W>
__global__ void __launch_bounds__(SPMV_THREADS, SPMV_MINBLOCKS) k_dmatmul_sten(const __grid_constant__ Sten st, SellView A, uint8_t bit, const uint8_t *__restrict__ ctl,
                                                                               double *__restrict__ x, const double *__restrict__ y, int pf_dist, int nsl)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = r >> 5;
  if (s >= nsl) return;                                         // whole warps
  const bool live = r < A.n && (!bit || (ctl[r] & bit));
  const long long cpo = __ldg(A.col_ptr + s);
  const int len = r < A.n ? (int)A.rowlen[r] : 0;
  if (cpo == st.code && __all_sync(0xffffffffu, len == W)) {
    const double xo = (OP != 0 && live) ? x[r] : 0.0;
    const char *yb = reinterpret_cast<const char *>(y + r);
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < W; j++) {
      const double yv = __ldg(reinterpret_cast<const double *>(yb + st.dbytes[j]));
      const double p = st.v[j] * yv;
      sum += p;
    }
    if (live) x[r] = OP == 0 ? sum : (OP == 1 ? xo + sum : xo - sum);
  } else {
    double s1[1], d1[1];
    row_product<1>(A, r, live, y, s1, d1);
    if (live) x[r] = OP == 0 ? s1[0] : (OP == 1 ? x[r] + s1[0] : x[r] - s1[0]);
  }
  const int lane = threadIdx.x & 31;
  if (pf_dist > 0 && s + pf_dist < nsl && lane < 5) {           // L2 prefetch: the far slice's rows of x and the rows of y it reaches first
    const size_t far = ((size_t)(s + pf_dist)) * 32;
    if (lane < 2) { if (OP != 0 && far + lane * 16 < (size_t)A.n) prefetch_l2(x + far + lane * 16); }
    else if (far + st.maxd + (lane - 2) * 16 < (size_t)A.n) prefetch_l2(y + far + st.maxd + (lane - 2) * 16);
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
  {
    // matrices with a dominant stencil, large levels: stencil rows and exception rows as two kernels (stx.cu); their byte model differs
    const double mb = stx_matrix_bytes(L, A);
    int done = 0;
    if (mb >= 0) {
      ProfScope ps(ctx, UGGPU_K_DMATMUL, (int)(L - ctx->lev), mb + (op == 0 ? 2.0 : 3.0) * nb);
      UG_TRY(stx_dmatmul(ctx, L, const_cast<SellMat *>(A), op, bit, x, y, &done));
    } else UG_TRY(stx_dmatmul(ctx, L, const_cast<SellMat *>(A), op, bit, x, y, &done));      // first use: builds the row mask (not timed)
    if (done) return 0;
  }
  ProfScope ps(ctx, UGGPU_K_DMATMUL, (int)(L - ctx->lev), A->entry_bytes() + 4.0 * (L->n + 1.0) + (op == 0 ? 2.0 : 3.0) * nb);
  const Prefetch pf = make_prefetch(ctx, A, BS);
  if (BS == 1 && (A->sten.w == 15 || A->sten.w == 27) && A->col_ptr != A->slice_ptr && !getenv("UGGPU_NO_STENCIL")) {
    const int nsl = (L->n + 31) / 32;
#define DS(OPV, WV) k_dmatmul_sten<OPV, WV><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(A->sten, v, bit, L->ctl, x, y, pf.dist, nsl)
    if (A->sten.w == 15) { if (op == 0) DS(0, 15); else if (op == 1) DS(1, 15); else DS(2, 15); }
    else { if (op == 0) DS(0, 27); else if (op == 1) DS(1, 27); else DS(2, 27); }
#undef DS
    KCHECK(ctx);
    return 0;
  }
  if (op == 0) k_dmatmul_k<BS, 0><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y, pf);
  else if (op == 1) k_dmatmul_k<BS, 1><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y, pf);
  else k_dmatmul_k<BS, 2><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(v, bit, L->ctl, x, y, pf);
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
__global__ void __launch_bounds__(SPMV_THREADS) k_jac_k(SellView A, const uint8_t *__restrict__ vclass, double *__restrict__ v, const double *__restrict__ d, Damp damp, int *err,
                                                        Prefetch pf, HaloK hk)
{
  constexpr int BB = BS * BS;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t cf = 0;
  if (hk.flag) {                                        // multi-GPU: this launch pushes v's interface rows into the neighbours' ghost rows
    halo_publish(hk);
    if ((r & ~31) < A.n) cf = hk.flag[r >> 5];          // used at the end only: the row's own loads do not wait for it
  }
  if ((r & ~31) >= A.n) return;                         // whole warps from here on
  const bool live = r < A.n;
  const int lane = r & 31;
  if (live && pf.dist > 0 && (r >> 5) + pf.dist < pf.nsl) {     // every stream of this kernel is direct-indexed: touch the far slice's lines now
    const PfState far{0, -1, (r >> 5) + pf.dist};
    if (lane < 2 * BB) prefetch_l2(reinterpret_cast<const char *>(A.diag) + ((size_t)far.slice * 32 * BB) * sizeof(double) + (size_t)lane * 128);
    pf_vec<BS>(d, far, pf);
    pf_rows<1>(vclass, far, pf);
  }
  double sol[BS];
  bool ok = live;
  if (live) {
    if (vclass[r] < 3) {
#pragma unroll
      for (int i = 0; i < BS; i++) sol[i] = 0.0;
    } else {
      const double *__restrict__ vp = A.diag + ((size_t)(r >> 5) * BB) * 32 + lane;
      double m[BB], rhs[BS];
#pragma unroll
      for (int k = 0; k < BB; k++) m[k] = vp[(size_t)k * 32];
#pragma unroll
      for (int i = 0; i < BS; i++) rhs[i] = d[(size_t)r * BS + i];
      if (solve_small_block<BS>(m, rhs, sol)) { atomicExch(err, UGGPU_SMALL_DIAG); ok = false; }
    }
  }
  const bool push = (cf & 2) && hk.peer;
  if (push) halo_wait(hk);                              // the neighbours' reads of the ghost rows about to be overwritten are done
  if (!ok) return;
  double pv[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) { pv[i] = sol[i] * damp.a[i]; v[(size_t)r * BS + i] = pv[i]; }
  if (push) halo_push_row<BS>(hk, r, pv);
}

// Timing experiment on ONE GPU (UGGPU_DBG_FAKE_COMM = 1: all-zero flags; 2: every 16th slice flagged "ghost columns + rows to push", waits
// and stores switched off): the comm instantiations of the kernels without a second GPU.  Results are unchanged.
__global__ void k_fake_flags(size_t n, int mode, uint8_t *f) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) f[i] = (mode >= 2 && (i & 15) == 15) ? 3 : 0; }
static int fake_comm(uggpu_ctx *ctx, Level *L, HaloK *hk)
{
  static uint8_t *flags = nullptr; static size_t cap = 0; static int built_mode = -1;
  static unsigned long long *go = nullptr;
  const int mode = atoi(getenv("UGGPU_DBG_FAKE_COMM"));
  const size_t need = ((size_t)L->n + 31) / 32 + 1;
  if (!go) { CUDA_TRY(cudaMalloc(&go, HALO_GO_SLOTS * 128)); CUDA_TRY(cudaMemset(go, 0, HALO_GO_SLOTS * 128)); }
  if (need > cap || mode != built_mode) {
    if (need > cap) { if (flags) cudaFree(flags); CUDA_TRY(cudaMalloc(&flags, need)); cap = need; }
    k_fake_flags<<<(int)((cap + 255) / 256), 256, 0, ctx->stream>>>(cap, mode, flags);
    built_mode = mode;
  }
  hk->flag = flags; hk->err = ctx->derr; hk->go = go; hk->sel = HALO_DBG_NOPUSH | HALO_DBG_NOWAIT | HALO_PUSH_B;
  hk->peer = mode >= 2 ? reinterpret_cast<double *const *>(go) : nullptr;
  return 0;
}

int k_jac(uggpu_ctx *ctx, int level, int A, double *v, const double *d, Damp damp, const HaloPlan *hp)
{
  Level *L = get_level(ctx, level);
  SellMat *M = get_mat(ctx, level, A);
  if (!L || !M) return UGGPU_DESC_MISMATCH;
  HaloK hk = halo_none();
  if (hp && hp->push) UG_TRY(halo_prepare(ctx, level, -1, M, nullptr, hp, &hk));
  if (hk.flag && getenv("UGGPU_DBG_HALO")) hk.sel |= atoi(getenv("UGGPU_DBG_HALO"));
  if (!hk.flag && getenv("UGGPU_DBG_FAKE_COMM")) UG_TRY(fake_comm(ctx, L, &hk));
  if (L->n == 0 && !hk.flag) return 0;
  int blocks = (L->n + SPMV_THREADS - 1) / SPMV_THREADS;
  if (blocks < 1) blocks = 1;
  SellView vw = view(*M);
  ProfScope ps(ctx, UGGPU_K_JAC, level, (double)L->n * (8.0 * L->bs * L->bs + 16.0 * L->bs));
  switch (L->bs) {
    case 1: k_jac_k<1><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr, make_prefetch(ctx, M, L->bs), hk); break;
    case 2: k_jac_k<2><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr, make_prefetch(ctx, M, L->bs), hk); break;
    default: k_jac_k<3><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(vw, L->vclass, v, d, damp, ctx->derr, make_prefetch(ctx, M, L->bs), hk); break;
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
// scalar rows: at most 32 registers, so that 2048 threads are resident per SM (the variant with the norm partials took 40 without
// the bound: 75 % occupancy, 4.69 instead of ~4.2 ms on the finest level)
// A slice of a partitioned level that reads ghost columns or holds rows to push (multi-GPU, HaloK): wait for the neighbours, the row
// product with L1-bypassing gathers where ghost columns occur, the step's updates, the push.  Same arithmetic as the kernels' own
// rows.  Kept out of line so that the registers of the smoothing kernels are those of their fast paths.
template <int BS, int FLAGS>
__device__ __noinline__ void smooth_comm_rows(SellView A, int r, int cf, HaloK hk, const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl,
                                              const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int *err, double *nrm)
{
  halo_wait(hk);
  const bool active = r < A.n;
  double s[BS], dg[BS * BS];
  if (cf & 1) row_product<BS, true>(A, r, active, tin, s, dg);
  else row_product<BS>(A, r, active, tin, s, dg);
#pragma unroll
  for (int i = 0; i < BS; i++) nrm[i] = 0.0;
  if (!active) return;
  double bn[BS], pv[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) {
    const size_t k = (size_t)r * BS + i;
    bn[i] = b[k] - s[i];
    b[k] = bn[i];
    pv[i] = bn[i];
  }
  if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
#pragma unroll
    for (int i = 0; i < BS; i++) {
      const size_t k = (size_t)r * BS + i;
      double cn;
      if (FLAGS & SF_CADD) cn = c[k] + tin[k];
      else if (FLAGS & SF_CSET) cn = 0.0 + tin[k];
      else cn = c[k];
      if (FLAGS & (SF_CADD | SF_CSET)) c[k] = cn;
      if (FLAGS & SF_XADD) x[k] = x[k] + cn;
      if ((hk.sel & 255) == HALO_PUSH_C) pv[i] = cn;
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
    for (int i = 0; i < BS; i++) {
      const double tv = sol[i] * damp.a[i];
      tout[(size_t)r * BS + i] = tv;
      if ((hk.sel & 255) == HALO_PUSH_TOUT) pv[i] = tv;
    }
  }
  if ((cf & 2) && hk.peer) halo_push_row<BS>(hk, r, pv);
  if (FLAGS & SF_NORM) {
    if (ctl[r] & UGGPU_CTL_NEW_DEFECT) {
#pragma unroll
      for (int i = 0; i < BS; i++) nrm[i] = bn[i] * bn[i];
    }
  }
}

template <int BS, int FLAGS, bool COMM = false>
__global__ void __launch_bounds__(SPMV_THREADS, BS == 1 ? (COMM ? SPMV_MINBLOCKS * 3 / 4 : SPMV_MINBLOCKS) : 1) k_smooth_k(SellView A, const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl,
                                                           const double *__restrict__ tin, double *__restrict__ b, double *__restrict__ c,
                                                           double *__restrict__ tout, Damp damp, double *__restrict__ x, double *__restrict__ partials, int *err,
                                                           Prefetch pf, const int32_t *__restrict__ list, int nlist, const uint8_t *__restrict__ skip_slice, HaloK hk)
{
  // Multi-GPU overlap (launch_smooth2): the INTERIOR launch covers the whole grid and skips the slices flagged in skip_slice (a
  // direct-indexed byte per slice: no extra hop in the row's chain of loads); the INTERFACE launch works on the slices list[w].
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (list) { const int w = r >> 5; r = w < nlist ? list[w] * 32 + (threadIdx.x & 31) : A.n + 32; }
  if (skip_slice && (r & ~31) < A.n && skip_slice[r >> 5]) r = A.n + 32;
  bool active = r < A.n;
  // Multi-GPU, peer-memory ghost rows (HaloK, uggpu_internal.h): block 0 tells the neighbours that this kernel has started; the warps
  // whose slice reads ghost columns or holds rows to push wait until all neighbours have started it, too
  // (COMM instantiation; the other one is the single-GPU kernel unchanged)
  uint8_t cf = 0;
  if (COMM && hk.flag) {
    halo_publish(hk);
    if ((r & ~31) < A.n) cf = hk.flag[r >> 5];
  }
  const PfState pfs = pf_begin(A, r, pf);      // software prefetch into L2 (uggpu_internal.h): requested now, issued at the end
  int64_t h_cpo = 0; int h_len = 0;            // COMM: the row's first loads leave together with the flag byte, not behind the branch on it
  if (COMM && (r & ~31) < A.n) row_head(A, r, active, h_cpo, h_len);
  double nrm[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) nrm[i] = 0.0;
  if (COMM && cf) {
    double gn[BS];
    smooth_comm_rows<BS, FLAGS>(A, r, cf, hk, vclass, ctl, tin, b, c, tout, damp, x, err, gn);
#pragma unroll
    for (int i = 0; i < BS; i++) nrm[i] = gn[i];
    active = false;                            // done
  }
  double s[BS], dg[BS * BS];
  // scalar rows: the row's own entries of b, c, tin are requested first, so that their (HBM or L2) round trip runs next to the gathers
  double eb = 0.0, ec = 0.0, et = 0.0;
  if (SPMV_EARLY && BS == 1 && active) {
    eb = b[r];
    if ((FLAGS & (SF_CADD | SF_XADD)) && !(FLAGS & SF_CSET)) ec = c[r];
    if (FLAGS & (SF_CADD | SF_CSET)) et = tin[r];
  }
  if (COMM) { if (!cf && (r & ~31) < A.n) row_product_h<BS>(A, r, h_cpo, h_len, tin, s, dg); }
  else if ((r & ~31) < A.n) row_product<BS>(A, r, active, tin, s, dg);
  if (active) {
    double bn[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) {
      size_t k = (size_t)r * BS + i;
      bn[i] = ((SPMV_EARLY && BS == 1) ? eb : b[k]) - s[i];
      b[k] = bn[i];
    }
    if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
#pragma unroll
      for (int i = 0; i < BS; i++) {
        size_t k = (size_t)r * BS + i;
        double cn;
        const double tk = (SPMV_EARLY && BS == 1) ? et : ((FLAGS & (SF_CADD | SF_CSET)) ? tin[k] : 0.0);
        const double ck = (SPMV_EARLY && BS == 1) ? ec : ((FLAGS & (SF_CADD | SF_XADD)) && !(FLAGS & SF_CSET) ? c[k] : 0.0);
        if (FLAGS & SF_CADD) cn = ck + tk;
        else if (FLAGS & SF_CSET) cn = 0.0 + tk;
        else cn = ck;
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
  pf_end<BS * BS>(A, pfs, pf);
  if ((pf.mode & 4) && pfs.sp >= 0) {
    pf_vec<BS>(b, pfs, pf);
    if (FLAGS & SF_CADD) pf_vec<BS>(c, pfs, pf);
    if (FLAGS & SF_TOUT) pf_rows<1>(vclass, pfs, pf);
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

// ---- fused smoothing step, stencil variant (scalar rows) -------------------------------------------------------------------
// Same arithmetic as k_smooth_k.  On a matrix whose slices mostly carry ONE stencil (Sten: the dominant pair of column-distance
// and value tables, sell_share_values) the thread-per-row kernel is no longer limited by HBM but by instruction issue: ncu on the
// 513^3 level showed ~490 warp instructions per slice, 2/3 of them bookkeeping (generic prefetch, 64-bit slice arithmetic, the
// paths for other storage forms).  Here the stencil arrives as a kernel parameter and its width W is a template parameter (15:
// P1 on simplices, 27: Q1 on hexahedra), so the loop over its columns is unrolled without predicates, every distance and value
// is a constant-bank operand (no table load, no shuffle), the row's own b, c, tin entries are requested before the gathers, and
// the prefetch touches only what such a slice streams (b, c, and the rows of tin first reached through the largest distance).
// Slices with another code word, or with rows shorter than the stencil, take row_product.
template <int FLAGS>
__device__ __forceinline__ double smooth_tail(int r, double sum, double dg, double eb, double ec, double et, uint8_t vc, const uint8_t *__restrict__ ctl,
                                              double *__restrict__ b, double *__restrict__ c, double *__restrict__ tout, double damp, double *__restrict__ x)
{
  const double bn = eb - sum;
  b[r] = bn;
  if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
    double cn;
    if (FLAGS & SF_CADD) cn = ec + et;
    else if (FLAGS & SF_CSET) cn = 0.0 + et;
    else cn = ec;
    if (FLAGS & (SF_CADD | SF_CSET)) c[r] = cn;
    if (FLAGS & SF_XADD) x[r] = x[r] + cn;
  }
  if (FLAGS & SF_TOUT) {
    const double sol = vc < 3 ? 0.0 : bn / dg;                 // l_jac: 0 below ACTIVE_CLASS (ugiter.cc:300)
    tout[r] = sol * damp;
  }
  if (FLAGS & SF_NORM) {
    if (ctl[r] & UGGPU_CTL_NEW_DEFECT) return bn * bn;
  }
  return 0.0;
}

template <int FLAGS, int W, bool COMM = false>
__global__ void __launch_bounds__(SPMV_THREADS, SPMV_MINBLOCKS) k_smooth_sten(const __grid_constant__ Sten st, SellView A, const uint8_t *__restrict__ vclass,
                                                                              const uint8_t *__restrict__ ctl, const double *__restrict__ tin, double *__restrict__ b,
                                                                              double *__restrict__ c, double *__restrict__ tout, double damp, double *__restrict__ x,
                                                                              double *__restrict__ partials, int *err, int pf_dist, int nsl, HaloK hk)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = r >> 5;
  const bool active = r < A.n;
  double nrm = 0.0;
  if (COMM && hk.flag) halo_publish(hk);
  if (s < nsl) {                                               // whole warps
    const long long cpo = __ldg(A.col_ptr + s);
    const int len = active ? (int)A.rowlen[r] : 0;
    // multi-GPU (HaloK, COMM instantiation): slices that read ghost columns or hold rows to push wait for the neighbours and take the careful path
    const uint8_t cf = (COMM && hk.flag) ? hk.flag[s] : (uint8_t)0;
    if (COMM && cf) {
      double gn[1];
      Damp dd; dd.a[0] = damp; dd.a[1] = dd.a[2] = 1.0;
      smooth_comm_rows<1, FLAGS>(A, r, cf, hk, vclass, ctl, tin, b, c, tout, dd, x, err, gn);
      nrm = gn[0];
    } else if (cpo == st.code && __all_sync(0xffffffffu, len == W)) {
      // the row's own entries first: their round trip runs next to the gathers
      const double eb = b[r];
      const double ec = ((FLAGS & (SF_CADD | SF_XADD)) && !(FLAGS & SF_CSET)) ? c[r] : 0.0;
      const double et = (FLAGS & (SF_CADD | SF_CSET)) ? tin[r] : 0.0;
      const uint8_t vc = (FLAGS & SF_TOUT) ? vclass[r] : (uint8_t)3;
      const char *yb = reinterpret_cast<const char *>(tin + r);
      double sum = 0.0;
#pragma unroll
      for (int j = 0; j < W; j++) {
        const double y = __ldg(reinterpret_cast<const double *>(yb + st.dbytes[j]));
        const double p = st.v[j] * y;
        sum += p;
      }
      nrm = smooth_tail<FLAGS>(r, sum, st.v[0], eb, ec, et, vc, ctl, b, c, tout, damp, x);
    } else {
      double s1[1], d1[1];
      row_product<1>(A, r, active, tin, s1, d1);
      if (active) {
        const double eb = b[r];
        const double ec = ((FLAGS & (SF_CADD | SF_XADD)) && !(FLAGS & SF_CSET)) ? c[r] : 0.0;
        const double et = (FLAGS & (SF_CADD | SF_CSET)) ? tin[r] : 0.0;
        const uint8_t vc = (FLAGS & SF_TOUT) ? vclass[r] : (uint8_t)3;
        nrm = smooth_tail<FLAGS>(r, s1[0], d1[0], eb, ec, et, vc, ctl, b, c, tout, damp, x);
      }
    }
    // L2 prefetch for the slice pf_dist ahead: its rows of b and c, and the rows of tin that slice reaches first (largest distance)
    const int lane = threadIdx.x & 31;
    if (pf_dist > 0 && s + pf_dist < nsl && lane < 8) {
      const size_t far = ((size_t)(s + pf_dist)) * 32;
      if (lane < 2) { if (far + lane * 16 < (size_t)A.n) prefetch_l2(b + far + lane * 16); }
      else if (lane < 4) { if (((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) && far + (lane - 2) * 16 < (size_t)A.n) prefetch_l2(c + far + (lane - 2) * 16); }
      else if (lane < 7) { if (far + st.maxd + (lane - 4) * 16 < (size_t)A.n) prefetch_l2(tin + far + st.maxd + (lane - 4) * 16); }
      else if (FLAGS & SF_TOUT) prefetch_l2(vclass + far);
    }
  }
  if (FLAGS & SF_NORM) {
    __shared__ double sm[SPMV_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double v = nrm;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
      v = lane < SPMV_THREADS / 32 ? sm[lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) partials[blockIdx.x] = v;
    }
  }
}

// ... and for 3x3 blocks (Q1 hexahedra, 27 block columns): the generic kernel needs 156 registers for block rows (19 % occupancy),
// which was fine while every thread streamed 2 KB of values per row, but leaves the SM idle once the values come from a shared
// table.  With the 27 distances and 243 values as constant-bank operands a row is 81 gathers + 243 multiply-adds in straight-line
// code.  Same sums in the same order as row_product_t (per entry and component i: acc = m_i0*w_0; acc += m_i1*w_1; acc += m_i2*w_2;
// s_i += acc), same epilogue as k_smooth_k.
#ifndef SPMV_STEN3_MINBLOCKS
#define SPMV_STEN3_MINBLOCKS 8
#endif
template <int FLAGS, bool COMM = false>
__global__ void __launch_bounds__(SPMV_THREADS, SPMV_STEN3_MINBLOCKS) k_smooth_sten3(const __grid_constant__ Sten3 st, SellView A, const uint8_t *__restrict__ vclass,
                                                                                     const uint8_t *__restrict__ ctl, const double *__restrict__ tin, double *__restrict__ b,
                                                                                     double *__restrict__ c, double *__restrict__ tout, Damp damp, double *__restrict__ x,
                                                                                     double *__restrict__ partials, int *err, int pf_dist, int nsl, HaloK hk)
{
  constexpr int BS = 3, BB = 9;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = r >> 5;
  const bool active = r < A.n;
  double nrm[BS] = {0.0, 0.0, 0.0};
  if (COMM && hk.flag) halo_publish(hk);
  if (s < nsl) {                                               // whole warps
    const long long cpo = __ldg(A.col_ptr + s);
    const int len = active ? (int)A.rowlen[r] : 0;
    double sum[BS], dg[BB];
    const uint8_t cf = (COMM && hk.flag) ? hk.flag[s] : (uint8_t)0;      // multi-GPU (HaloK): ghost columns / rows to push in this slice
    bool work = active;
    if (COMM && cf) {
      double gn[BS];
      smooth_comm_rows<BS, FLAGS>(A, r, cf, hk, vclass, ctl, tin, b, c, tout, damp, x, err, gn);
#pragma unroll
      for (int i = 0; i < BS; i++) nrm[i] = gn[i];
      work = false;
    } else if (cpo == st.code && __all_sync(0xffffffffu, len == 27)) {
      const char *yb = reinterpret_cast<const char *>(tin + (size_t)r * BS);
#pragma unroll
      for (int i = 0; i < BS; i++) sum[i] = 0.0;
#pragma unroll
      for (int k = 0; k < BB; k++) dg[k] = st.v[k];
#pragma unroll
      for (int j = 0; j < 27; j++) {
        const double *yp = reinterpret_cast<const double *>(yb + st.dbytes[j]);
        const double w0 = __ldg(yp), w1 = __ldg(yp + 1), w2 = __ldg(yp + 2);
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double acc = st.v[j * BB + i * BS] * w0;
          acc = acc + st.v[j * BB + i * BS + 1] * w1;
          acc = acc + st.v[j * BB + i * BS + 2] * w2;
          sum[i] += acc;
        }
      }
    } else {
      row_product<BS>(A, r, active, tin, sum, dg);
    }
    if (work) {
      double bn[BS];
#pragma unroll
      for (int i = 0; i < BS; i++) {
        const size_t k = (size_t)r * BS + i;
        bn[i] = b[k] - sum[i];
        b[k] = bn[i];
      }
      if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
#pragma unroll
        for (int i = 0; i < BS; i++) {
          const size_t k = (size_t)r * BS + i;
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
    // L2 prefetch for the slice pf_dist ahead: its rows of b and c (6 lines each) and the rows of tin first reached through the largest distance
    const int lane = threadIdx.x & 31;
    if (pf_dist > 0 && s + pf_dist < nsl && lane < 19) {
      const size_t far = ((size_t)(s + pf_dist)) * 32 * BS;
      const size_t nn = (size_t)A.n * BS;
      if (lane < 6) { if (far + lane * 16 < nn) prefetch_l2(b + far + lane * 16); }
      else if (lane < 12) { if (((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) && far + (lane - 6) * 16 < nn) prefetch_l2(c + far + (lane - 6) * 16); }
      else { const size_t o = far + (size_t)st.maxd * BS + (lane - 12) * 16; if (o < nn) prefetch_l2(tin + o); }
    }
  }
  if (FLAGS & SF_NORM) {
    __shared__ double sm[SPMV_THREADS / 32][UGGPU_MAX_BS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
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

// ---- fused smoothing step, asynchronously staged through shared memory ----------------------------------------------
// Same arithmetic as k_smooth_k, scalar rows.  The thread-per-row kernel above keeps every byte of the matrix stream in
// registers while it is in flight and runs out of outstanding loads long before it runs out of HBM bandwidth (ncu: 59 %
// DRAM throughput, 0.35 eligible warps per scheduler, everything waiting on L1TEX).  Here every warp owns two rings of
// shared-memory buffers and works TWO slices ahead of the arithmetic:
//   values    one cp.async.bulk (TMA) per slice -- SELL-32 keeps a slice's values contiguous -- completing on an mbarrier;
//   operand   the gathered entries of the correction: one 8-byte cp.async per lane and entry, straight from L2/HBM into
//             shared memory (no register is held while the load is in flight), tracked by commit groups.
// The arithmetic then reads shared memory only.  One persistent CTA per SM; warp w of CTA c takes slices
// (c + k * gridDim.x) * warps + w, k = 0, 1, ...
#define TMA_VSTAGES 3
#define TMA_XSTAGES 3
#define TMA_MAX_WARPS 16
#define TMA_SMEM_BUDGET (216 * 1024)
#define TMA_PREFETCH 8
#define TMA_MAX_WIDTH 48

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int FLAGS>
__global__ void __launch_bounds__(TMA_MAX_WARPS * 32, 1) k_smooth_tma(SellView A, const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl,
                                                                       const double *__restrict__ tin, double *__restrict__ b, double *__restrict__ c,
                                                                       double *__restrict__ tout, double damp, double *__restrict__ x,
                                                                       double *__restrict__ partials, int *err, int maxw)
{
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5, wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsl = (A.n + 31) >> 5;
  // ring slots (doubles): values of one slice; gathered operand of one slice + 4 rows of per-row vector entries (b, c, t, x)
  const int vslot = maxw * 32, xslot = (maxw + 4) * 32;
  double *vring = reinterpret_cast<double *>(smem) + (size_t)wi * (TMA_VSTAGES * vslot + TMA_XSTAGES * xslot);
  double *xring = vring + TMA_VSTAGES * vslot;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)warps * (TMA_VSTAGES * vslot + TMA_XSTAGES * xslot) * sizeof(double)) + wi * TMA_VSTAGES;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < TMA_VSTAGES; i++) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const int stride = (int)gridDim.x * warps;                    // in slices
  const int s_first = (int)blockIdx.x * warps + wi;
  const int nit = s_first < nsl ? (nsl - 1 - s_first) / stride + 1 : 0;
  const uint32_t vbase = smem_u32(vring), xbase = smem_u32(xring) + (uint32_t)lane * 8u;

  // "far" registers: offsets and row length of the next slice to stage, loaded one step before they are used
  int64_t f_sp0 = 0, f_sp1 = 0, f_cpo = 0;
  int f_len = 0;
  int64_t d_cpo = 0;      // distance table held in dtab (0: none; uniform slices have col_ptr < 0)
  int dtab[32];
  auto load_far = [&](int sl) {
    if (sl < nsl) {
      f_sp0 = __ldg(A.slice_ptr + sl); f_sp1 = __ldg(A.slice_ptr + sl + 1); f_cpo = __ldg(A.col_ptr + sl);
      const int r = sl * 32 + lane;
      f_len = r < A.n ? (int)A.rowlen[r] : 0;
    }
  };
  // stage slice sl (k-th of this warp): bulk copy of its values, cp.async of its operand entries and of the row's vector
  // entries; returns the row length
  auto stage = [&](int k, int sl) -> int {
    int len = 0;
    if (sl < nsl) {
      len = f_len;
      const int w = (int)((f_sp1 - f_sp0) >> 5);
      const int vs = k % TMA_VSTAGES, xs = k % TMA_XSTAGES;
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[vs], (uint32_t)w * 256u);
        if (w > 0) bulk_g2s(vring + vs * vslot, A.val + f_sp0, (uint32_t)w * 256u, &bars[vs]);
      }
      const int r = sl * 32 + lane;
      const uint32_t xdst = xbase + (uint32_t)(xs * xslot) * 8u;
      const double *__restrict__ trow = tin + r;
      if (f_cpo < 0) {
        if (f_cpo != d_cpo) {                 // new distance table: one coalesced load, then kept in registers (slices share tables)
          const int dreg = __ldg(A.col + UG_COLTAB(f_cpo) + lane);      // the column array ends with 32 spare words
#pragma unroll
          for (int j = 0; j < 32; j++) dtab[j] = __shfl_sync(0xffffffffu, dreg, j);
          d_cpo = f_cpo;
        }
#pragma unroll
        for (int j = 0; j < 16; j++)
          if (j < len) cp_async8(xdst + (uint32_t)j * 256u, trow + dtab[j]);
        if (w > 16) {
#pragma unroll
          for (int j = 16; j < 32; j++)
            if (j < len) cp_async8(xdst + (uint32_t)j * 256u, trow + dtab[j]);
        }
      } else {
        const int32_t *__restrict__ cp = A.col + f_cpo + lane;
#pragma unroll 8
        for (int j = 0; j < len; j++) cp_async8(xdst + (uint32_t)j * 256u, tin + __ldg(cp + (size_t)j * 32));
      }
      if (r < A.n) {
        const uint32_t rdst = xdst + (uint32_t)maxw * 256u;
        cp_async8(rdst, b + r);
        if ((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) cp_async8(rdst + 256u, c + r);
        if (FLAGS & (SF_CADD | SF_CSET)) cp_async8(rdst + 512u, trow);
        if (FLAGS & SF_XADD) cp_async8(rdst + 768u, x + r);
      }
    }
    cp_async_commit();            // one group per call, empty or not: cp_async_wait<2> below counts calls
    return len;
  };

  // the small per-slice / per-row streams (slice offsets, row lengths, flags) are pulled into L2 TMA_PREFETCH slices ahead by
  // five lanes of one prefetch instruction, so that the loads that need them never wait for HBM
  auto prefetch_meta = [&](int slp) {
    if (slp < nsl && lane < 5) {
      const void *p = lane == 0 ? (const void *)(A.slice_ptr + slp + 1) : lane == 1 ? (const void *)(A.col_ptr + slp)
                    : lane == 2 ? (const void *)(A.rowlen + (size_t)slp * 32) : lane == 3 ? (const void *)(vclass + (size_t)slp * 32)
                                                                                          : (const void *)(ctl + (size_t)slp * 32);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  };
#pragma unroll 1
  for (int k = 0; k < TMA_PREFETCH; k++) prefetch_meta(s_first + k * stride);
  int len0, len1, len2 = 0;
  load_far(s_first); len0 = stage(0, s_first);
  load_far(s_first + stride); len1 = stage(1, s_first + stride);
  load_far(s_first + 2 * stride);
  double nacc = 0.0;
  int sl = s_first;
#pragma unroll 1
  for (int it = 0; it < nit; it++, sl += stride) {
    prefetch_meta(sl + TMA_PREFETCH * stride);
    len2 = stage(it + 2, sl + 2 * stride);
    load_far(sl + 3 * stride);
    const int r = sl * 32 + lane;
    const bool live = r < A.n;
    int vc = 3, cb = 0;
    if (live) {
      if (FLAGS & SF_TOUT) vc = vclass[r];
      if (FLAGS & SF_NORM) cb = ctl[r];
    }
    const int vs = it % TMA_VSTAGES, xs = it % TMA_XSTAGES;
    cp_async_wait<2>();
    while (!mbar_try_wait(&bars[vs], (uint32_t)((it / TMA_VSTAGES) & 1))) { }
    const double *sval = vring + vs * vslot + lane;
    const double *sx = xring + xs * xslot + lane;
    double s = 0.0;
    if (__all_sync(0xffffffffu, len0 <= 32)) {
#pragma unroll
      for (int j = 0; j < 16; j++)
        if (j < len0) s += sval[j * 32] * sx[j * 32];
      if (__any_sync(0xffffffffu, len0 > 16)) {
#pragma unroll
        for (int j = 16; j < 32; j++)
          if (j < len0) s += sval[j * 32] * sx[j * 32];
      }
    } else {
#pragma unroll 8
      for (int j = 0; j < len0; j++) s += sval[j * 32] * sx[j * 32];
    }
    const double dg = sval[0];
    const double *srow = sx + maxw * 32;
    double bv = 0.0, cv = 0.0, tv = 0.0, xv = 0.0;
    if (live) {
      bv = srow[0];
      if ((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) cv = srow[32];
      if (FLAGS & (SF_CADD | SF_CSET)) tv = srow[64];
      if (FLAGS & SF_XADD) xv = srow[96];
    }
    __syncwarp();
    if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ring reads done before the next bulk write
    if (live) {
      const double bn = bv - s;
      b[r] = bn;
      if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
        double cn;
        if (FLAGS & SF_CADD) cn = cv + tv;
        else if (FLAGS & SF_CSET) cn = 0.0 + tv;
        else cn = cv;
        if (FLAGS & (SF_CADD | SF_CSET)) c[r] = cn;
        if (FLAGS & SF_XADD) x[r] = xv + cn;
      }
      if (FLAGS & SF_TOUT) {
        const double sol = vc < 3 ? 0.0 : bn / dg;
        tout[r] = sol * damp;
      }
      if (FLAGS & SF_NORM) { if (cb & UGGPU_CTL_NEW_DEFECT) nacc += bn * bn; }
    }
    len0 = len1; len1 = len2;
  }
  cp_async_wait<0>();
  if (FLAGS & SF_NORM) {
    for (int o = 16; o > 0; o >>= 1) nacc += __shfl_down_sync(0xffffffffu, nacc, o);
    if (lane == 0) partials[(size_t)blockIdx.x * warps + wi] = nacc;
  }
}

// ---- multi-GPU overlap: interior and interface slices ----------------------------------------------------------------------
// flag[s] = 1 when a row of slice s has an entry in a ghost column (column index >= number of owned rows)
__global__ void k_slice_has_ghost(SellView A, int n_owned, uint8_t *__restrict__ flag)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if ((r & ~31) >= A.n) return;
  bool ghost = false;
  if (r < A.n) {
    const int len = A.rowlen[r];
    const ColIter ci = col_iter(A, r);
    for (int j = 0; j < len; j++) if (col_at(ci, j) >= n_owned) ghost = true;
  }
  ghost = __any_sync(0xffffffffu, ghost);
  if ((threadIdx.x & 31) == 0) flag[r >> 5] = ghost ? 1 : 0;
}

static int build_overlap_lists(uggpu_ctx *ctx, Level *L, SellMat *A)
{
  const size_t nsl = (size_t)(A->n + 31) / 32;
  uint8_t *d_flag = nullptr;
  UG_TRY(dalloc(ctx, &d_flag, nsl));
  k_slice_has_ghost<<<(int)((nsl * 32 + 255) / 256), 256, 0, ctx->stream>>>(view(*A), L->n, d_flag);
  KCHECK(ctx);
  std::vector<uint8_t> flag(nsl);
  CUDA_TRY(cudaMemcpyAsync(flag.data(), d_flag, nsl, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  A->bnd_flag = d_flag;
  std::vector<int32_t> bd;
  for (size_t s = 0; s < nsl; s++) if (flag[s]) bd.push_back((int32_t)s);
  A->n_int = (int)(nsl - bd.size()); A->n_bnd = (int)bd.size();
  UG_TRY(dalloc(ctx, &A->bnd_list, (size_t)(A->n_bnd > 0 ? A->n_bnd : 1)));
  if (A->n_bnd) CUDA_TRY(cudaMemcpyAsync(A->bnd_list, bd.data(), sizeof(int32_t) * bd.size(), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// geometry of the staged kernel for matrix A; false: use the thread-per-row kernel
static bool tma_geometry(uggpu_ctx *ctx, const Level *L, const SellMat *A, int *grid, int *warps, size_t *smem)
{
  // environment switches (tests, A/B runs): UGGPU_NO_TMA, UGGPU_TMA_MIN_ROWS (default 32768: smaller levels are launch-bound anyway)
  if (!getenv("UGGPU_TMA")) return false;        // opt-in: measured slower than the prefetching thread-per-row kernel
  const char *mr = getenv("UGGPU_TMA_MIN_ROWS");
  const int minrows = mr ? atoi(mr) : (1 << 15);
  if (L->bs != 1 || A->maxlen < 1 || A->maxlen > TMA_MAX_WIDTH || L->n < minrows) return false;
  const size_t per_warp = ((size_t)TMA_VSTAGES * A->maxlen + (size_t)TMA_XSTAGES * (A->maxlen + 4)) * 256 + TMA_VSTAGES * sizeof(uint64_t);
  int w = (int)((TMA_SMEM_BUDGET - 256) / per_warp);
  const char *mw = getenv("UGGPU_TMA_WARPS");
  if (mw && atoi(mw) > 0 && atoi(mw) < w) w = atoi(mw);
  if (w > TMA_MAX_WARPS) w = TMA_MAX_WARPS;
  if (w < 2) return false;
  const int64_t nsl = ((int64_t)L->n + 31) / 32;
  int64_t g = (nsl + w - 1) / w;
  if (g > ctx->sm_count) g = ctx->sm_count;
  *grid = (int)g; *warps = w;
  *smem = (size_t)w * per_warp;
  return true;
}

// split: the halo exchange of tin has only been STARTED (halo_begin): the interior slices run on the compute stream right away, the
// interface slices on the halo stream behind the wait + unpack; the compute stream continues when both are done.  Every row is
// computed by the same code on the same operands as in one launch: results do not change.
template <int BS, int FLAGS>
static int launch_smooth2(uggpu_ctx *ctx, Level *L, SellMat *A, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot,
                          int level, int split, const HaloK &hk)
{
  int blocks = (L->n + SPMV_THREADS - 1) / SPMV_THREADS;
  if (blocks < 1) blocks = 1;
  int tgrid = 0, twarps = 0; size_t tsmem = 0;
  const bool tma = !split && !hk.flag && BS == 1 && tma_geometry(ctx, L, A, &tgrid, &twarps, &tsmem);
  constexpr int WPB = SPMV_THREADS / 32;
  int bi = 0, bb = 0;
  if (split) {
    if (A->n_int < 0) UG_TRY(build_overlap_lists(ctx, L, A));
    bi = (L->n + SPMV_THREADS - 1) / SPMV_THREADS; bb = (A->n_bnd + WPB - 1) / WPB;     // interior: the whole grid minus the flagged slices
    blocks = bi + bb;
  }
  if (tma) blocks = tgrid * twarps;     // one partial per warp
  if (FLAGS & SF_NORM) UG_TRY(ensure_partials(ctx, (size_t)blocks * BS));
  const double nb = 8.0 * BS * L->n;
  // algorithmic bytes: entries, row lengths, gathered operand once, b read+write, c, tout, x
  ProfScope ps(ctx, UGGPU_K_SMOOTH, (int)(L - ctx->lev), A->entry_bytes() + 4.0 * (L->n + 1.0) + 3.0 * nb
               + ((FLAGS & SF_CADD) ? 2.0 * nb : 0.0) + ((FLAGS & SF_CSET) ? nb : 0.0) + ((FLAGS & SF_TOUT) ? nb : 0.0) + ((FLAGS & SF_XADD) ? 2.0 * nb : 0.0));
  if (tma) {
    static bool attr_set = false;     // per instantiation
    if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(k_smooth_tma<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BUDGET)); attr_set = true; }
    k_smooth_tma<FLAGS><<<tgrid, twarps * 32, tsmem, ctx->stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp.a[0], x, ctx->partials, ctx->derr, A->maxlen);
  } else if (split) {
    if (!ctx->halo_stream) {
      int lo = 0, hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_TRY(cudaStreamCreateWithPriority(&ctx->halo_stream, cudaStreamNonBlocking, hi));      // its blocks go first when SM slots free up
      for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventCreateWithFlags(&ctx->halo_ev[i], cudaEventDisableTiming));
    }
    const Prefetch pf = make_prefetch(ctx, A, BS);
    CUDA_TRY(cudaEventRecord(ctx->halo_ev[0], ctx->stream));                  // behind the push of halo_begin
    if (bi > 0) {
      k_smooth_k<BS, FLAGS><<<bi, SPMV_THREADS, 0, ctx->stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr, pf, nullptr, 0, A->bnd_flag, halo_none());
      KCHECK(ctx);
    }
    CUDA_TRY(cudaStreamWaitEvent(ctx->halo_stream, ctx->halo_ev[0], 0));
    UG_TRY(halo_finish(ctx, level, const_cast<double *>(tin), ctx->halo_stream));
    if (bb > 0) {
      k_smooth_k<BS, FLAGS><<<bb, SPMV_THREADS, 0, ctx->halo_stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x,
                                                                        (FLAGS & SF_NORM) ? ctx->partials + (size_t)bi * BS : ctx->partials, ctx->derr, pf, A->bnd_list, A->n_bnd, nullptr, halo_none());
      ctx->launches++;
    }
    CUDA_TRY(cudaEventRecord(ctx->halo_ev[1], ctx->halo_stream));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->halo_ev[1], 0));
  } else if (BS == 1 && (A->sten.w == 15 || A->sten.w == 27) && A->col_ptr != A->slice_ptr && !getenv("UGGPU_NO_STENCIL")) {
    // most slices carry one stencil: the variant with the stencil in the constant bank (same arithmetic, a third of the instructions)
    const Prefetch pf = make_prefetch(ctx, A, BS);
#define STEN_LAUNCH(WV, CV) k_smooth_sten<FLAGS, WV, CV><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(A->sten, view(*A), L->vclass, L->ctl, tin, b, c, tout, damp.a[0], x, \
                                                                                                   ctx->partials, ctx->derr, pf.dist, (L->n + 31) / 32, hk)
    if (A->sten.w == 15) { if (hk.flag) STEN_LAUNCH(15, true); else STEN_LAUNCH(15, false); }
    else { if (hk.flag) STEN_LAUNCH(27, true); else STEN_LAUNCH(27, false); }
#undef STEN_LAUNCH
  } else if (BS == 3 && A->sten3 && A->col_ptr != A->slice_ptr && !getenv("UGGPU_NO_STENCIL")) {
    // the same for 3x3 blocks with the 27-column stencil of Q1 hexahedra (pf distance as for vector-only slices)
    const Prefetch pf = make_prefetch(ctx, A, BS);
    if (hk.flag)
      k_smooth_sten3<FLAGS, true><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(*A->sten3, view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr,
                                                                           pf.dist, (L->n + 31) / 32, hk);
    else
      k_smooth_sten3<FLAGS, false><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(*A->sten3, view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr,
                                                                            pf.dist, (L->n + 31) / 32, hk);
  } else {
    if (hk.flag)
      k_smooth_k<BS, FLAGS, true><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr,
                                                                              make_prefetch(ctx, A, BS), nullptr, 0, nullptr, hk);
    else
      k_smooth_k<BS, FLAGS, false><<<blocks, SPMV_THREADS, 0, ctx->stream>>>(view(*A), L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr,
                                                                               make_prefetch(ctx, A, BS), nullptr, 0, nullptr, hk);
  }
  KCHECK(ctx);
  if (FLAGS & SF_NORM) UG_TRY(reduce_partials_final(ctx, BS, (size_t)blocks, norm_slot, (int)(L - ctx->lev)));
  return 0;
}

template <int BS>
static int launch_smooth(uggpu_ctx *ctx, Level *L, SellMat *A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot,
                         int level, int split, const HaloK &hk)
{
#define SM_CASE(F) case F: return launch_smooth2<BS, F>(ctx, L, A, tin, b, c, tout, damp, x, norm_slot, level, split, hk)
  switch (flags) {
    SM_CASE(0);
    SM_CASE(SF_CADD);
    SM_CASE(SF_CSET);
    SM_CASE(SF_TOUT);
    SM_CASE(SF_CADD | SF_TOUT);
    SM_CASE(SF_CSET | SF_TOUT);
    SM_CASE(SF_CADD | SF_XADD | SF_NORM);
    SM_CASE(SF_CSET | SF_XADD | SF_NORM);
    SM_CASE(SF_CADD | SF_XADD | SF_NORM | SF_TOUT);      // last step of a cycle that another cycle follows (cycle.cu TopFuse::want_t)
    SM_CASE(SF_CSET | SF_XADD | SF_NORM | SF_TOUT);
    SM_CASE(SF_CADD | SF_XADD);
    SM_CASE(SF_CSET | SF_XADD);
    SM_CASE(SF_CADD | SF_NORM);
    SM_CASE(SF_CSET | SF_NORM);
    SM_CASE(SF_NORM);
  }
#undef SM_CASE
  return uggpu_fail(UGGPU_ERROR, "smooth step: unsupported flag combination %d", flags);
}

int k_smooth_step(uggpu_ctx *ctx, int level, int A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot,
                  const HaloPlan *hp)
{
  Level *L = get_level(ctx, level);
  SellMat *M = get_mat(ctx, level, A);
  if (!L || !M) return UGGPU_DESC_MISMATCH;
  if ((flags & SF_TOUT) && tout == tin) return uggpu_fail(UGGPU_ERROR, "smooth step: tout aliases tin");
  // Ghost columns of the correction (no-op on one GPU).  Peer-memory ghost rows: the kernel itself waits for the neighbours and stores the
  // interface rows of the vector it produces into their ghost rows (hk); window transport with UGGPU_OVERLAP: only the push happens here
  // and the kernel's interior slices overlap the rest of the exchange; otherwise the operand is exchanged before the launch.
  int split = 0;
  HaloK hk = halo_none();
  if (halo_fused_available(ctx, level)) {
    UG_TRY(halo_prepare(ctx, level, level, M, const_cast<double *>(tin), hp, &hk));
    if (hk.peer) {
      hk.sel = hp->push == tout ? HALO_PUSH_TOUT : (hp->push == b ? HALO_PUSH_B : (hp->push == c ? HALO_PUSH_C : HALO_PUSH_NONE));
      if (hk.sel == HALO_PUSH_NONE || ((hk.sel & 255) == HALO_PUSH_TOUT && !(flags & SF_TOUT)) || ((hk.sel & 255) == HALO_PUSH_C && !(flags & (SF_CADD | SF_CSET))))
        return uggpu_fail(UGGPU_ERROR, "smooth step: the vector to push is not produced by this step");
    }
  } else UG_TRY(halo_begin(ctx, level, const_cast<double *>(tin), &split));
  if (!hk.flag && getenv("UGGPU_DBG_FAKE_COMM")) UG_TRY(fake_comm(ctx, L, &hk));
  if (L->n == 0 && !hk.flag) return 0;
  if (!split) {       // matrices with a dominant stencil, large levels: stencil rows and exception rows as two kernels (stx.cu)
    int done = 0;
    UG_TRY(stx_smooth(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot, hk, &done));
    if (done) return 0;
  }
  switch (L->bs) {
    case 1: return launch_smooth<1>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot, level, split, hk);
    case 2: return launch_smooth<2>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot, level, split, hk);
    default: return launch_smooth<3>(ctx, L, M, flags, tin, b, c, tout, damp, x, norm_slot, level, split, hk);
  }
}
