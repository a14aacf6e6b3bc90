// stx.cu -- "stencil rows + exception rows": the SpMV-type kernels of a matrix with a dominant stencil as TWO kernels.
//
// spmv.cu's stencil kernels (k_smooth_sten, k_dmatmul_sten) decide per SLICE: a slice whose 32 rows all carry the dominant stencil runs
// the unrolled constant-bank loop, every other slice -- one Dirichlet row is enough -- the generic row product, which lives four to
// five times as long (a chain of dependent loads: code word -> column words -> gathers).  On a uniformly refined grid those slices
// are few but not rare (7 % at 513^3, 28 % at 129^3; on the upper half of an x-split twice as many as on the lower one, which is what
// made the ranks of a partitioned run unequal), and being latency-bound they cost their share several times over.
//
// Here the decision is per ROW.  Bit l of xmask[s] says that row 32 s + l is EXACTLY the stencil: same length, same column distances,
// bit-identical values, no ghost column, nothing to push to another GPU.  Kernel 1 (k_*_stx) handles those rows -- one 4-byte mask per
// warp is all it reads of the matrix -- and nothing else: no code word, no row length, no slow path, no communication.  Kernel 2
// (k_*_xrows) takes the compact ascending list of all other rows, one thread per row with the generic product on the SELL arrays: 32
// exception rows per warp instead of one or two.  Both kernels perform, for every row, the same operations on the same operands in
// the same order as the one-kernel forms (canonical VSTART->MNEXT order): results are bit-identical (tests: UGGPU_NO_STX=1 A/B, port).
//
// Multi-GPU (peer-memory ghost rows, HaloK): rows with ghost columns and rows whose result a neighbour needs are exception rows, so
// kernel 1 never waits and never stores remotely; kernel 2 runs NEXT to it on a second stream, waits for the neighbours (one go word
// per SM), gathers ghost columns past L1, and pushes -- the interface rows of an x-split, one per grid line in the row order, sit in
// neighbouring lanes of the exception list, so their remote stores share sectors.
#include "uggpu_internal.h"

#include <algorithm>
#include <cstdlib>
#include <vector>

#define STX_THREADS 128
#define STX_MINBLOCKS (2048 / STX_THREADS)
#define STX_MIN_ROWS (1 << 20)      // smaller levels are launch-bound: one kernel (spmv.cu) is better than two

// ---- which rows are exactly the stencil ---------------------------------------------------------------------------------------------
template <int BS, class STEN>
__global__ void k_stx_rowmask(const __grid_constant__ STEN st, SellView A, int n_owned, const uint32_t *__restrict__ snd_bits, uint32_t *__restrict__ xmask)
{
  constexpr int BB = BS * BS;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if ((r & ~31) >= A.n) return;
  const int lane = r & 31;
  bool match = r < A.n && (int)A.rowlen[r < A.n ? r : 0] == st.w;
  if (match) {
    const int64_t sp = slice_off(A, r >> 5);
    const double *vp = A.val + sp * BB + lane;
    const ColIter ci = col_iter(A, r);
    for (int j = 0; j < st.w && match; j++) {
      const int c = col_at(ci, j);
      if ((long long)(c - r) * (long long)(BS * sizeof(double)) != st.dbytes[j] || c < 0 || c >= n_owned) match = false;
      for (int k = 0; k < BB && match; k++)
        if (__double_as_longlong(vp[((size_t)j * BB + k) * 32]) != __double_as_longlong(st.v[j * BB + k])) match = false;
    }
    if (snd_bits && ((snd_bits[r >> 5] >> lane) & 1u)) match = false;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, match);
  if (lane == 0) xmask[r >> 5] = m;
}

// copies the entries of the exception rows into the packed form (one thread per list position)
// Entries whose stored value (block) is all zeros -- the off-diagonal entries of a Dirichlet row, the diagonal neighbours of the P1 Laplacian
// on a Kuhn mesh -- are not copied into the packed rows, and the stencil kernels skip the stencil's zero coefficients (sten_nonzero).  The
// accumulator of a row product starts at +0.0 and the sum of two doubles is -0.0 only if both are, so it is never -0.0; a product with a
// zero coefficient is +-0.0 for every FINITE operand and adding it changes nothing: the results are bit-identical to the reference's as
// long as the operand holds no Inf / NaN (then the reference's row turns NaN through 0 * Inf and this one does not -- the solve is lost
// either way).  What it saves: a Dirichlet row gathered 15 operands, one 32-byte sector each, for 14 zero coefficients (690 B of DRAM
// traffic per row, 15 % of the traffic of a smoothing step at 513^3).  UGGPU_KEEP_ZERO_ENTRIES=1 keeps every entry (A/B).
static bool keep_zero_entries() { static int k = -1; if (k < 0) k = getenv("UGGPU_KEEP_ZERO_ENTRIES") ? 1 : 0; return k == 1; }

template <int BS>
__device__ __forceinline__ bool stx_entry_kept(const double *vp, int j, bool keep_all)
{
  if (keep_all || j == 0) return true;
  for (int k = 0; k < BS * BS; k++) if (vp[((size_t)j * BS * BS + k) * 32] != 0.0) return true;
  return false;
}

// entries the packed copy of exception row i keeps
template <int BS>
__global__ void k_stx_count(SellView A, const int32_t *__restrict__ xrows, int nx, int keep_all, uint16_t *__restrict__ cnt)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  const int r = xrows[i];
  const double *vp = A.val + slice_off(A, r >> 5) * BS * BS + (r & 31);
  const int len = (int)A.rowlen[r];
  int c = 0;
  for (int j = 0; j < len; j++) if (stx_entry_kept<BS>(vp, j, keep_all != 0)) c++;
  cnt[i] = (uint16_t)c;
}

template <int BS>
__global__ void k_stx_pack(SellView A, const int32_t *__restrict__ xrows, int nx, int keep_all, const int64_t *__restrict__ xs_ptr, uint16_t *__restrict__ xs_len,
                           int32_t *__restrict__ xs_col, double *__restrict__ xs_val)
{
  constexpr int BB = BS * BS;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  const int r = xrows[i], sl = r >> 5, lane = r & 31, pl = i & 31;
  const int64_t sp = slice_off(A, sl);
  const int64_t cpo = (A.fixed_w && A.col_ptr == A.slice_ptr) ? sp : A.col_ptr[sl];
  const int len = (int)A.rowlen[r];
  const bool uni = cpo < 0;
  const int32_t *cp = uni ? A.col + UG_COLTAB(cpo) : A.col + cpo + lane;
  const int cstride = uni ? 1 : 32, cbase = uni ? r : 0;
  const double *vp = A.val + sp * BB + lane;          // the explicit values are always complete
  const int64_t o = xs_ptr[i >> 5];
  int q = 0;
  for (int j = 0; j < len; j++) {
    if (!stx_entry_kept<BS>(vp, j, keep_all != 0)) continue;
    xs_col[o + (int64_t)q * 32 + pl] = cp[(size_t)j * cstride] + cbase;
    for (int k = 0; k < BB; k++) xs_val[(o + (int64_t)q * 32) * BB + (int64_t)k * 32 + pl] = vp[((size_t)j * BB + k) * 32];
    q++;
  }
  xs_len[i] = (uint16_t)q;
}

int stx_free(uggpu_ctx *ctx, SellMat *m)
{
  const size_t nsl = ((size_t)(m->n > 0 ? m->n : 0) + 31) / 32;
  const size_t nxs = ((size_t)(m->nx > 0 ? m->nx : 0) + 31) / 32;
  if (m->xs_ptr) dfree(ctx, m->xs_ptr, nxs + 1);
  if (m->xs_len) dfree(ctx, m->xs_len, nxs * 32 + 1);
  if (m->xs_col) dfree(ctx, m->xs_col, (size_t)m->xs_entries + 1);
  if (m->xs_val) dfree(ctx, m->xs_val, (size_t)m->xs_entries * m->bb + 1);
  m->xs_entries = 0;
  if (m->xmask) dfree(ctx, m->xmask, nsl + 1);
  if (m->xrows) dfree(ctx, m->xrows, (size_t)(m->nx > 0 ? m->nx : 0) + 1);
  m->nx = -1; m->x_comm = 0;
  return 0;
}

// builds xmask / xrows of A (rows of level L); comm: ghost columns and rows to push are exceptions (peer-memory ghost transport)
static int stx_ensure(uggpu_ctx *ctx, Level *L, SellMat *A, bool comm)
{
  if (A->xmask && A->x_comm == (comm ? 1 : 0)) return 0;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  UG_TRY(stx_free(ctx, A));
  const size_t nsl = ((size_t)A->n + 31) / 32;
  UG_TRY(dalloc(ctx, &A->xmask, nsl + 1));
  const uint32_t *snd = nullptr;
  if (comm) {
    if (!L->halo) return uggpu_fail(UGGPU_ERROR, "stx: the level's halo tables are not set up");
    snd = halo_snd_bits(L);      // rows to push (comm.cu)
  }
  const int blocks = (int)((nsl * 32 + 255) / 256);
  if (A->bb == 1) k_stx_rowmask<1, Sten><<<blocks, 256, 0, ctx->stream>>>(A->sten, view(*A), L->n, snd, A->xmask);
  else k_stx_rowmask<3, Sten3><<<blocks, 256, 0, ctx->stream>>>(*A->sten3, view(*A), L->n, snd, A->xmask);
  KCHECK(ctx);
  std::vector<uint32_t> mask(nsl);
  CUDA_TRY(cudaMemcpyAsync(mask.data(), A->xmask, sizeof(uint32_t) * nsl, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> rows;
  for (size_t s = 0; s < nsl; s++) {
    const uint32_t m = mask[s];
    if (m == 0xffffffffu) continue;
    for (int l = 0; l < 32; l++) {
      const int64_t r = (int64_t)s * 32 + l;
      if (r < A->n && !((m >> l) & 1u)) rows.push_back((int32_t)r);
    }
  }
  A->nx = (int)rows.size();
  UG_TRY(dalloc(ctx, &A->xrows, rows.size() + 1));
  if (!rows.empty()) CUDA_TRY(cudaMemcpyAsync(A->xrows, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, ctx->stream));
  // packed copy of the exception rows' entries: widths per group of 32 list positions from the number of entries each row keeps
  const int keep_all = keep_zero_entries() ? 1 : 0;
  std::vector<uint16_t> rl(rows.size());
  if (!rows.empty()) {
    uint16_t *d_cnt = nullptr;
    UG_TRY(dalloc(ctx, &d_cnt, rows.size()));
    const int cb = (A->nx + 255) / 256;
    if (A->bb == 1) k_stx_count<1><<<cb, 256, 0, ctx->stream>>>(view(*A), A->xrows, A->nx, keep_all, d_cnt);
    else k_stx_count<3><<<cb, 256, 0, ctx->stream>>>(view(*A), A->xrows, A->nx, keep_all, d_cnt);
    KCHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(rl.data(), d_cnt, sizeof(uint16_t) * rl.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    dfree(ctx, d_cnt, rows.size());
  }
  const size_t nxs = (rows.size() + 31) / 32;
  std::vector<int64_t> xp(nxs + 1, 0);
  for (size_t g = 0; g < nxs; g++) {
    int w = 0;
    for (size_t i = g * 32; i < rows.size() && i < g * 32 + 32; i++) w = std::max(w, (int)rl[i]);
    xp[g + 1] = xp[g] + (int64_t)w * 32;
  }
  A->xs_entries = xp[nxs];
  UG_TRY(dalloc(ctx, &A->xs_ptr, nxs + 1));
  UG_TRY(dalloc(ctx, &A->xs_len, nxs * 32 + 1));
  UG_TRY(dalloc(ctx, &A->xs_col, (size_t)A->xs_entries + 1));
  UG_TRY(dalloc(ctx, &A->xs_val, (size_t)A->xs_entries * A->bb + 1));
  CUDA_TRY(cudaMemcpyAsync(A->xs_ptr, xp.data(), sizeof(int64_t) * (nxs + 1), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaMemsetAsync(A->xs_len, 0, sizeof(uint16_t) * (nxs * 32 + 1), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(A->xs_col, 0, sizeof(int32_t) * ((size_t)A->xs_entries + 1), ctx->stream));
  CUDA_TRY(cudaMemsetAsync(A->xs_val, 0, sizeof(double) * ((size_t)A->xs_entries * A->bb + 1), ctx->stream));
  if (A->nx > 0) {
    const int pb = (A->nx + 255) / 256;
    if (A->bb == 1) k_stx_pack<1><<<pb, 256, 0, ctx->stream>>>(view(*A), A->xrows, A->nx, keep_all, A->xs_ptr, A->xs_len, A->xs_col, A->xs_val);
    else k_stx_pack<3><<<pb, 256, 0, ctx->stream>>>(view(*A), A->xrows, A->nx, keep_all, A->xs_ptr, A->xs_len, A->xs_col, A->xs_val);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  A->x_comm = comm ? 1 : 0;
  return 0;
}

// ---- the product of list position i on the packed copy of the exception rows (coalesced: entry j of the warp's 32 rows is contiguous) ----
// CG: ghost columns (index >= n_owned) are gathered past L1 (see spmv.cu gather_ld)
struct XPack { const int64_t *ptr; const uint16_t *len; const int32_t *col; const double *val; };
template <int BS, bool CG>
__device__ __forceinline__ void row_product_packed(const XPack &X, int i, int n_owned, const double *__restrict__ y, double (&s)[BS], double (&dg)[BS * BS])
{
  constexpr int BB = BS * BS;
  const int pl = i & 31;
  const int64_t o = __ldg(X.ptr + (i >> 5));
  const int len = (int)X.len[i];
  const int32_t *__restrict__ cp = X.col + o + pl;
  const double *__restrict__ vp = X.val + o * BB + pl;
#pragma unroll
  for (int q = 0; q < BS; q++) s[q] = 0.0;
#pragma unroll
  for (int k = 0; k < BB; k++) dg[k] = 0.0;
#pragma unroll 4
  for (int j = 0; j < len; j++) {
    const int c = __ldg(cp + (size_t)j * 32);
    double m[BB], w[BS];
#pragma unroll
    for (int k = 0; k < BB; k++) m[k] = __ldg(vp + ((size_t)j * BB + k) * 32);
#pragma unroll
    for (int q = 0; q < BS; q++) w[q] = (CG && c >= n_owned) ? __ldcg(y + (size_t)c * BS + q) : y[(size_t)c * BS + q];
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < BB; k++) dg[k] = m[k];
    }
#pragma unroll
    for (int q = 0; q < BS; q++) {
      double acc = m[q * BS] * w[0];
#pragma unroll
      for (int t = 1; t < BS; t++) acc = acc + m[q * BS + t] * w[t];
      s[q] += acc;
    }
  }
}

// the smoothing step's updates of one row (the same statements as spmv.cu k_smooth_k); returns the norm contributions
template <int BS, int FLAGS>
__device__ __forceinline__ void smooth_row_tail(int r, const double (&s)[BS], const double (&dg)[BS * BS], const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl,
                                                const double *__restrict__ tin, double *__restrict__ b, double *__restrict__ c, double *__restrict__ tout, const Damp &damp,
                                                double *__restrict__ x, int *err, int sel, double (&pv)[BS], double (&nrm)[BS])
{
  double bn[BS];
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
      if (FLAGS & SF_CADD) cn = c[k];
      else if (FLAGS & SF_CSET) cn = 0.0;
      else cn = c[k];
      if (FLAGS & SF_CPREV) cn = cn + tout[k];                 // the previous step's correction (uggpu_internal.h SF_CPREV)
      if (FLAGS & (SF_CADD | SF_CSET)) cn = cn + tin[k];
      if (FLAGS & (SF_CADD | SF_CSET)) c[k] = cn;
      if (FLAGS & SF_XADD) x[k] = x[k] + cn;
      if ((sel & 255) == HALO_PUSH_C) pv[i] = cn;
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
      if ((sel & 255) == HALO_PUSH_TOUT) pv[i] = tv;
    }
  }
  if (FLAGS & SF_NORM) {
    if (ctl[r] & UGGPU_CTL_NEW_DEFECT) {
#pragma unroll
      for (int i = 0; i < BS; i++) nrm[i] = bn[i] * bn[i];
    }
  }
}

template <int BS, int THREADS>
__device__ __forceinline__ void block_norm_partials(const double (&nrm)[BS], double *__restrict__ partials)
{
  __shared__ double sm[THREADS / 32][UGGPU_MAX_BS];
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
      double v = lane < THREADS / 32 ? sm[lane][i] : 0.0;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) partials[(size_t)blockIdx.x * BS + i] = v;
    }
  }
}

struct YTile { int q, k, block; };                            // slices between two passes, passes (k = 1: off); block: the k slices go to the k warps of a block instead of k passes of a warp

// ---- kernel 1, scalar rows: the rows that are exactly the stencil ---------------------------------------------------------------------
template <int FLAGS, int W, int YK>
__global__ void __launch_bounds__(STX_THREADS, STX_MINBLOCKS) k_smooth_stx(const __grid_constant__ Sten st, int n, const uint32_t *__restrict__ xmask,
                                                                            const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl, const double *__restrict__ tin,
                                                                            double *__restrict__ b, double *__restrict__ c, double *__restrict__ tout, double damp,
                                                                            double *__restrict__ x, double *__restrict__ partials, int pf_dist, int nsl, YTile yt)
{
  // Slice order.  YK == 1: warp w takes slice w.  YK > 1 ("y tiles"): a warp takes YK slices that are yt.q slices apart -- the same x
  // range of YK consecutive grid lines when yt.q slices make one line (stx_ytile reads it off the stencil's second distance) -- so that
  // the neighbouring lines a row's gathers touch are the ones its own warp fetched in the previous pass and still sit in L1: less L2 -> L1
  // traffic for the gathers (measured at 513^3, 7-entry stencil: pair 1.22 -> 1.13 ms with YK = 4; 2: 1.17, 8: 1.17).  Warp j of group g
  // takes slices g*q*YK + j + q*i, i < YK: a permutation of the slices, every slice exactly once, whatever the grid.  Only instantiated
  // for W = 7: the loop costs registers the 21- and 27-entry variants do not have under the 32-register bound (they spill: 1.58 -> 3.1 ms).
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // YK == 1 with yt.k == 4: the same permutation spread over the four warps of a block (block j of group g: slices g*q*4 + j + q*w) --
  // no loop, no extra registers: the form the 21- / 27-entry variants can afford
  const int g0 = YK == 1 ? (yt.block ? ((wg >> 2) / yt.q) * yt.q * 4 + (wg >> 2) % yt.q + yt.q * (wg & 3) : wg) : (wg / yt.q) * yt.q * YK + wg % yt.q;
  double nrm = 0.0;
#pragma unroll 1
  for (int it = 0; it < YK; it++) {
    const int s = YK == 1 ? g0 : g0 + it * yt.q;
    if (s >= nsl) break;                                       // whole warps
    const int r = s * 32 + lane;
    const uint32_t m = __ldg(xmask + s);
    if ((m >> lane) & 1u) {
      // the row's own entries first: their round trip runs next to the gathers
      const double eb = b[r];
      const double ec = ((FLAGS & (SF_CADD | SF_XADD)) && !(FLAGS & SF_CSET)) ? c[r] : 0.0;
      const double et = (FLAGS & (SF_CADD | SF_CSET)) ? tin[r] : 0.0;
      const double ep = (FLAGS & SF_CPREV) ? tout[r] : 0.0;
      const uint8_t vc = (FLAGS & SF_TOUT) ? vclass[r] : (uint8_t)3;
      const char *yb = reinterpret_cast<const char *>(tin + r);
      double sum = 0.0;
#pragma unroll
      for (int j = 0; j < W; j++) {
        const double y = __ldg(reinterpret_cast<const double *>(yb + st.dbytes[j]));
        const double p = st.v[j] * y;
        sum += p;
      }
      const double bn = eb - sum;
      b[r] = bn;
      if (FLAGS & (SF_CADD | SF_CSET | SF_XADD)) {
        double cn;
        if (FLAGS & SF_CADD) cn = (FLAGS & SF_CPREV) ? (ec + ep) + et : ec + et;
        else if (FLAGS & SF_CSET) cn = (FLAGS & SF_CPREV) ? (0.0 + ep) + et : 0.0 + et;
        else cn = ec;
        if (FLAGS & (SF_CADD | SF_CSET)) c[r] = cn;
        if (FLAGS & SF_XADD) x[r] = x[r] + cn;
      }
      if (FLAGS & SF_TOUT) {
        const double sol = vc < 3 ? 0.0 : bn / st.v[0];          // l_jac: 0 below ACTIVE_CLASS (ugiter.cc:300)
        tout[r] = sol * damp;
      }
      if (FLAGS & SF_NORM) { if (ctl[r] & UGGPU_CTL_NEW_DEFECT) nrm = nrm + bn * bn; }
    }
    // L2 prefetch for the slice pf_dist ahead: its rows of b and c, the rows of tin that slice reaches first (largest distance), its mask
    if (pf_dist > 0 && s + pf_dist < nsl && lane < 9) {
      const size_t far = ((size_t)(s + pf_dist)) * 32;
      if (lane < 2) { if (far + lane * 16 < (size_t)n) prefetch_l2(b + far + lane * 16); }
      else if (lane < 4) { if (((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) && far + (lane - 2) * 16 < (size_t)n) prefetch_l2(c + far + (lane - 2) * 16); }
      else if (lane < 7) { if (far + st.maxd + (lane - 4) * 16 < (size_t)n) prefetch_l2(tin + far + st.maxd + (lane - 4) * 16); }
      else if (lane == 7) { if (FLAGS & SF_TOUT) prefetch_l2(vclass + far); }
      else prefetch_l2(xmask + s + pf_dist);
    }
  }
  if (FLAGS & SF_NORM) { const double na[1] = {nrm}; block_norm_partials<1, STX_THREADS>(na, partials); }
}

// ... and 3x3 blocks (27 block columns)
#ifndef STX3_MINBLOCKS
#define STX3_MINBLOCKS 8
#endif
template <int FLAGS>
__global__ void __launch_bounds__(STX_THREADS, STX3_MINBLOCKS) k_smooth_stx3(const __grid_constant__ Sten3 st, int n, const uint32_t *__restrict__ xmask,
                                                                             const uint8_t *__restrict__ vclass, const uint8_t *__restrict__ ctl, const double *__restrict__ tin,
                                                                             double *__restrict__ b, double *__restrict__ c, double *__restrict__ tout, Damp damp,
                                                                             double *__restrict__ x, double *__restrict__ partials, int *err, int pf_dist, int nsl)
{
  constexpr int BS = 3, BB = 9;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = r >> 5, lane = threadIdx.x & 31;
  double nrm[BS] = {0.0, 0.0, 0.0};
  if (s < nsl) {
    const uint32_t m = __ldg(xmask + s);
    if ((m >> lane) & 1u) {
      double sum[BS], dg[BB], pv[BS];
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
      smooth_row_tail<BS, FLAGS>(r, sum, dg, vclass, ctl, tin, b, c, tout, damp, x, err, 0, pv, nrm);
    }
    if (pf_dist > 0 && s + pf_dist < nsl && lane < 20) {
      const size_t far = ((size_t)(s + pf_dist)) * 32 * BS;
      const size_t nn = (size_t)n * BS;
      if (lane < 6) { if (far + lane * 16 < nn) prefetch_l2(b + far + lane * 16); }
      else if (lane < 12) { if (((FLAGS & SF_CADD) || ((FLAGS & SF_XADD) && !(FLAGS & SF_CSET))) && far + (lane - 6) * 16 < nn) prefetch_l2(c + far + (lane - 6) * 16); }
      else if (lane < 19) { const size_t o = far + (size_t)st.maxd * BS + (lane - 12) * 16; if (o < nn) prefetch_l2(tin + o); }
      else prefetch_l2(xmask + s + pf_dist);
    }
  }
  if (FLAGS & SF_NORM) block_norm_partials<BS, STX_THREADS>(nrm, partials);
}

// ---- kernel 2: the exception rows, one thread per row ----------------------------------------------------------------------------------
// COMM (multi-GPU, HaloK): block 0 publishes / watches, every warp waits for this SM's go word before it gathers ghost columns or pushes
template <int BS, int FLAGS, bool COMM>
__global__ void __launch_bounds__(STX_THREADS) k_smooth_xrows(XPack X, int n_owned, const int32_t *__restrict__ xrows, int nx, const uint8_t *__restrict__ vclass,
                                                              const uint8_t *__restrict__ ctl, const double *__restrict__ tin, double *__restrict__ b, double *__restrict__ c,
                                                              double *__restrict__ tout, Damp damp, double *__restrict__ x, double *__restrict__ partials, int *err, HaloK hk)
{
  if (COMM) { halo_publish(hk); halo_wait(hk); }
  double nrm[BS];
#pragma unroll
  for (int q = 0; q < BS; q++) nrm[q] = 0.0;
  // one thread per row, or -- with a capped grid (stx_xgrid) -- a grid-stride loop: a narrow kernel that runs UNDER the stencil rows
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    const int r = __ldg(xrows + i);
    double s[BS], dg[BS * BS], pv[BS];
    row_product_packed<BS, COMM>(X, i, n_owned, tin, s, dg);
    smooth_row_tail<BS, FLAGS>(r, s, dg, vclass, ctl, tin, b, c, tout, damp, x, err, COMM ? hk.sel : 0, pv, nrm);
    if (COMM && hk.peer) halo_push_row<BS>(hk, r, pv);
  }
  if (FLAGS & SF_NORM) block_norm_partials<BS, STX_THREADS>(nrm, partials);
}

// ---- dmatmul family ---------------------------------------------------------------------------------------------------------------------
template <int OP, int W, int YK>
__global__ void __launch_bounds__(STX_THREADS, STX_MINBLOCKS) k_dmatmul_stx(const __grid_constant__ Sten st, int n, const uint32_t *__restrict__ xmask, uint8_t bit,
                                                                             const uint8_t *__restrict__ ctl, double *__restrict__ x, const double *__restrict__ y, int pf_dist, int nsl,
                                                                             YTile yt)
{
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int g0 = YK == 1 ? wg : (wg / yt.q) * yt.q * YK + wg % yt.q;      // slice order: see k_smooth_stx
#pragma unroll 1
  for (int it = 0; it < YK; it++) {
  const int s = YK == 1 ? wg : g0 + it * yt.q;
  if (s >= nsl) return;
  const int r = s * 32 + lane;
  const uint32_t m = __ldg(xmask + s);
  if (((m >> lane) & 1u) && (!bit || (ctl[r] & bit))) {
    const double xo = OP != 0 ? x[r] : 0.0;
    const char *yb = reinterpret_cast<const char *>(y + r);
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < W; j++) {
      const double yv = __ldg(reinterpret_cast<const double *>(yb + st.dbytes[j]));
      const double p = st.v[j] * yv;
      sum += p;
    }
    x[r] = OP == 0 ? sum : (OP == 1 ? xo + sum : xo - sum);
  }
  if (pf_dist > 0 && s + pf_dist < nsl && lane < 6) {
    const size_t far = ((size_t)(s + pf_dist)) * 32;
    if (lane < 2) { if (OP != 0 && far + lane * 16 < (size_t)n) prefetch_l2(x + far + lane * 16); }
    else if (lane < 5) { if (far + st.maxd + (lane - 2) * 16 < (size_t)n) prefetch_l2(y + far + st.maxd + (lane - 2) * 16); }
    else prefetch_l2(xmask + s + pf_dist);
  }
  }
}

template <int BS, int OP>
__global__ void __launch_bounds__(STX_THREADS) k_dmatmul_xrows(XPack X, int n_owned, const int32_t *__restrict__ xrows, int nx, uint8_t bit, const uint8_t *__restrict__ ctl,
                                                               double *__restrict__ x, const double *__restrict__ y)
{
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += gridDim.x * blockDim.x) {
    const int r = __ldg(xrows + i);
    if (bit && !(ctl[r] & bit)) continue;
    double s[BS], dg[BS * BS];
    row_product_packed<BS, false>(X, i, n_owned, y, s, dg);
#pragma unroll
    for (int q = 0; q < BS; q++) {
      const size_t k = (size_t)r * BS + q;
      if (OP == 0) x[k] = (BS == 1) ? s[q] : 0.0 + s[q];
      else if (OP == 1) x[k] = x[k] + s[q];
      else x[k] = x[k] - s[q];
    }
  }
}

// ---- launches ---------------------------------------------------------------------------------------------------------------------------
static bool stx_applies(const Level *L, const SellMat *A)
{
  if (getenv("UGGPU_NO_STX") || getenv("UGGPU_NO_STENCIL")) return false;
  const char *mr = getenv("UGGPU_STX_MIN_ROWS");
  if (L->n < (mr ? atoi(mr) : STX_MIN_ROWS)) return false;
  if (A->col_ptr == A->slice_ptr) return false;
  if (L->bs == 1) return A->bb == 1 && (A->sten.w == 15 || A->sten.w == 27);
  return L->bs == 3 && A->bb == 9 && A->sten3 != nullptr;
}

bool stx_handles(const Level *L, const SellMat *A) { return stx_applies(L, A) && !getenv("UGGPU_NO_CPREV"); }

// bytes of MATRIX data one pass of the kernel pair fetches: the row mask, and the packed copy of the exception rows (values, columns, list,
// lengths) -- the stencil rows read nothing else of the matrix.  < 0: the pair does not apply to this matrix (yet)
double stx_matrix_bytes(const Level *L, const SellMat *A)
{
  if (!stx_applies(L, A) || !A->xmask) return -1.0;
  return 4.0 * ((L->n + 31) / 32) + (double)A->xs_entries * (8.0 * A->bb + 4.0) + 6.0 * (double)(A->nx > 0 ? A->nx : 0);
}

// Grid of an exception-row kernel.  Launched with one thread per row and high priority it takes the whole GPU for its duration (a chain
// of dependent loads per row: 0.2 ms at 513^3 with the SMs nearly idle) and the stencil rows start behind it.  With UGGPU_XROWS_CTAS = k
// CTAs per SM it is a narrow grid-stride kernel instead, resident next to the stencil-row kernel for that kernel's whole duration.
static int stx_xgrid(const uggpu_ctx *ctx, int xblocks)
{
  static int k = -1;
  if (k < 0) { const char *e = getenv("UGGPU_XROWS_CTAS"); k = e ? atoi(e) : 0; }
  if (k <= 0) return xblocks;
  const int cap = ctx->sm_count * k;
  return xblocks < cap ? xblocks : cap;
}

static int stx_fork(uggpu_ctx *ctx, cudaStream_t *xs)
{
  if (!ctx->halo_stream) {
    int lo = 0, hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&ctx->halo_stream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; i++) CUDA_TRY(cudaEventCreateWithFlags(&ctx->halo_ev[i], cudaEventDisableTiming));
  }
  *xs = ctx->halo_stream;
  CUDA_TRY(cudaEventRecord(ctx->halo_ev[0], ctx->stream));
  CUDA_TRY(cudaStreamWaitEvent(*xs, ctx->halo_ev[0], 0));
  return 0;
}
static int stx_join(uggpu_ctx *ctx, cudaStream_t xs)
{
  CUDA_TRY(cudaEventRecord(ctx->halo_ev[1], xs));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->halo_ev[1], 0));
  return 0;
}

// the stencil without its zero coefficients (entry 0, the diagonal, stays first); *w_out = number of entries kept
static Sten sten_nonzero(const Sten &st)
{
  if (keep_zero_entries()) return st;
  Sten c = st;
  int q = 0;
  for (int j = 0; j < st.w; j++)
    if (j == 0 || st.v[j] != 0.0) { c.dbytes[q] = st.dbytes[j]; c.v[q] = st.v[j]; q++; }
  for (int j = q; j < 32; j++) { c.dbytes[j] = 0; c.v[j] = 0.0; }
  c.w = q;
  if (q != 7 && q != 15 && q != 21 && q != 27) return st;      // widths the kernels are instantiated for (P1 Kuhn: 15 -> 7, Q1 Laplace: 27 -> 21)
  return c;
}

// y tiles of the scalar stencil kernel (k_smooth_stx): passes per warp from UGGPU_STX_YTILE (default below), slices between two passes
// from the stencil's second-smallest positive distance (the grid line: N for the 7-point, N - 1 for the 27-point stencil)
#ifndef STX_YTILE_DEFAULT
#define STX_YTILE_DEFAULT 4
#endif
static YTile stx_ytile(const Sten &st, int nsl)
{
  static int k = -1;
  if (k < 0) { const char *e = getenv("UGGPU_STX_YTILE"); k = e ? atoi(e) : STX_YTILE_DEFAULT; if (k < 1) k = 1; }
  YTile yt{1, 1, 0};
  static int bt = -1;
  if (bt < 0) bt = getenv("UGGPU_STX_BTILE") ? 1 : 0;
  if (k != 4 || (st.w != 7 && !bt)) return yt;                   // W = 7: four passes per warp; other widths (UGGPU_STX_BTILE=1): four warps of a block
  long long best = 0;
  for (int j = 0; j < st.w; j++) {
    const long long d = st.dbytes[j] / (long long)sizeof(double);
    if (d > 1 && (best == 0 || d < best)) best = d;
  }
  const int q = (int)((best + 16) / 32);
  if (q < 1 || (long long)q * k > nsl) return yt;
  yt.q = q; yt.k = k; yt.block = st.w != 7;
  return yt;
}
static int stx_yblocks(const YTile &yt, int nsl)
{
  const long long groups = ((long long)nsl + (long long)yt.q * yt.k - 1) / ((long long)yt.q * yt.k);
  if (yt.block) return (int)(groups * yt.q);                     // one block (k = 4 warps = STX_THREADS / 32) per k slices
  return (int)((groups * yt.q * 32 + STX_THREADS - 1) / STX_THREADS);
}

template <int BS, int FLAGS>
static int stx_smooth2(uggpu_ctx *ctx, Level *L, SellMat *A, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot, const HaloK &hk)
{
  const int nsl = (L->n + 31) / 32;
  const Sten stc = BS == 1 ? sten_nonzero(A->sten) : Sten();
  const YTile yt = BS == 1 ? stx_ytile(stc, nsl) : YTile{1, 1, 0};
  const int blocks = BS == 1 ? stx_yblocks(yt, nsl) : (L->n + STX_THREADS - 1) / STX_THREADS;
  const int xblocks = stx_xgrid(ctx, (A->nx + STX_THREADS - 1) / STX_THREADS > 0 ? (A->nx + STX_THREADS - 1) / STX_THREADS : 1);
  if (FLAGS & SF_NORM) UG_TRY(ensure_partials(ctx, (size_t)(blocks + xblocks) * BS));
  const double nb = 8.0 * BS * L->n;
  // algorithmic bytes: what the pair reads of the matrix (mask + packed exception rows), the gathered operand once, b read + write, c, tout, x
  ProfScope ps(ctx, UGGPU_K_SMOOTH, (int)(L - ctx->lev), stx_matrix_bytes(L, A) + 3.0 * nb
               + ((FLAGS & SF_CADD) ? 2.0 * nb : 0.0) + ((FLAGS & SF_CSET) ? nb : 0.0) + ((FLAGS & SF_TOUT) ? nb : 0.0) + ((FLAGS & SF_XADD) ? 2.0 * nb : 0.0)
               + ((FLAGS & SF_CPREV) ? nb : 0.0));
  Prefetch pf = make_prefetch(ctx, A, BS);
  // the exception rows run NEXT to the stencil rows on a second stream: a small latency-bound kernel (1-2 % of the rows, a chain of
  // dependent loads per row) that would otherwise leave the GPU half empty for its whole duration -- and, multi-GPU, the one that waits
  const bool side = hk.flag || !getenv("UGGPU_STX_SAME_STREAM");
  cudaStream_t xs = ctx->stream;
  if (side) UG_TRY(stx_fork(ctx, &xs));
  double *xpart = ctx->partials + (size_t)blocks * BS;
  const XPack X{A->xs_ptr, A->xs_len, A->xs_col, A->xs_val};
  if (hk.flag) k_smooth_xrows<BS, FLAGS, true><<<xblocks, STX_THREADS, 0, xs>>>(X, L->n, A->xrows, A->nx, L->vclass, L->ctl, tin, b, c, tout, damp, x, xpart, ctx->derr, hk);
  else if (A->nx > 0 || (FLAGS & SF_NORM)) k_smooth_xrows<BS, FLAGS, false><<<xblocks, STX_THREADS, 0, xs>>>(X, L->n, A->xrows, A->nx, L->vclass, L->ctl, tin, b, c, tout, damp, x, xpart, ctx->derr, hk);
  KCHECK(ctx);
  if (BS == 1) {
    const Sten &st = stc;
#define SX(WV, YKV) k_smooth_stx<FLAGS, WV, YKV><<<blocks, STX_THREADS, 0, ctx->stream>>>(st, L->n, A->xmask, L->vclass, L->ctl, tin, b, c, tout, damp.a[0], x, ctx->partials, pf.dist, nsl, yt)
    if (st.w == 7) { if (yt.k == 4) SX(7, 4); else SX(7, 1); } else if (st.w == 15) SX(15, 1); else if (st.w == 21) SX(21, 1); else SX(27, 1);
#undef SX
  } else {
    k_smooth_stx3<FLAGS><<<blocks, STX_THREADS, 0, ctx->stream>>>(*A->sten3, L->n, A->xmask, L->vclass, L->ctl, tin, b, c, tout, damp, x, ctx->partials, ctx->derr, pf.dist, nsl);
  }
  KCHECK(ctx);
  if (side) UG_TRY(stx_join(ctx, xs));
  if (FLAGS & SF_NORM) UG_TRY(reduce_partials_final(ctx, BS, (size_t)(blocks + xblocks), norm_slot, (int)(L - ctx->lev)));
  return 0;
}

template <int BS>
static int stx_smooth1(uggpu_ctx *ctx, Level *L, SellMat *A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot, const HaloK &hk)
{
#define SM_CASE(F) case F: return stx_smooth2<BS, F>(ctx, L, A, tin, b, c, tout, damp, x, norm_slot, hk)
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
    // the second step of a pair whose first step left c alone (SF_CPREV)
    SM_CASE(SF_CADD | SF_CPREV);
    SM_CASE(SF_CSET | SF_CPREV);
    SM_CASE(SF_CADD | SF_CPREV | SF_TOUT);
    SM_CASE(SF_CSET | SF_CPREV | SF_TOUT);
    SM_CASE(SF_CADD | SF_CPREV | SF_XADD | SF_NORM);
    SM_CASE(SF_CADD | SF_CPREV | SF_XADD | SF_NORM | SF_TOUT);
    SM_CASE(SF_CADD | SF_CPREV | SF_XADD);
    SM_CASE(SF_CADD | SF_CPREV | SF_NORM);
    SM_CASE(SF_CSET | SF_CPREV | SF_XADD | SF_NORM);
    SM_CASE(SF_CSET | SF_CPREV | SF_XADD | SF_NORM | SF_TOUT);
    SM_CASE(SF_CSET | SF_CPREV | SF_XADD);
    SM_CASE(SF_CSET | SF_CPREV | SF_NORM);
  }
#undef SM_CASE
  return uggpu_fail(UGGPU_ERROR, "smooth step: unsupported flag combination %d", flags);
}

int stx_smooth(uggpu_ctx *ctx, Level *L, SellMat *A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot,
               const HaloK &hk, int *done)
{
  *done = 0;
  if (!stx_applies(L, A)) return 0;
  UG_TRY(stx_ensure(ctx, L, A, hk.flag != nullptr));
  *done = 1;
  if (L->bs == 1) return stx_smooth1<1>(ctx, L, A, flags, tin, b, c, tout, damp, x, norm_slot, hk);
  return stx_smooth1<3>(ctx, L, A, flags, tin, b, c, tout, damp, x, norm_slot, hk);
}

int stx_dmatmul(uggpu_ctx *ctx, Level *L, SellMat *A, int op, uint8_t bit, double *x, const double *y, int *done)
{
  *done = 0;
  if (L->bs != 1 || !stx_applies(L, A)) return 0;
  // the mask built for the comm form (ghost columns and rows to push are exceptions) serves as well: exception rows are computed generically
  if (!A->xmask) UG_TRY(stx_ensure(ctx, L, A, false));
  *done = 1;
  const int nsl = (L->n + 31) / 32;
  const int xfull = (A->nx + STX_THREADS - 1) / STX_THREADS, xblocks = stx_xgrid(ctx, xfull);
  const Prefetch pf = make_prefetch(ctx, A, 1);
  // a capped exception-row grid runs next to the stencil rows on the second stream (started first); otherwise behind them on the same stream
  const bool side = xblocks > 0 && xblocks < xfull;
  cudaStream_t xs = ctx->stream;
  if (side) {
    UG_TRY(stx_fork(ctx, &xs));
    const XPack X{A->xs_ptr, A->xs_len, A->xs_col, A->xs_val};
    if (op == 0) k_dmatmul_xrows<1, 0><<<xblocks, STX_THREADS, 0, xs>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    else if (op == 1) k_dmatmul_xrows<1, 1><<<xblocks, STX_THREADS, 0, xs>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    else k_dmatmul_xrows<1, 2><<<xblocks, STX_THREADS, 0, xs>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    KCHECK(ctx);
  }
  const Sten st = sten_nonzero(A->sten);
  YTile yt = stx_ytile(st, nsl);
  if (yt.block) yt = YTile{1, 1, 0};                              // the block form exists in the smoothing kernel only
  const int sblocks = stx_yblocks(yt, nsl);
#define DS(OPV, WV, YKV) k_dmatmul_stx<OPV, WV, YKV><<<sblocks, STX_THREADS, 0, ctx->stream>>>(st, L->n, A->xmask, bit, L->ctl, x, y, pf.dist, nsl, yt)
#define DW(WV, YKV) { if (op == 0) DS(0, WV, YKV); else if (op == 1) DS(1, WV, YKV); else DS(2, WV, YKV); }
  if (st.w == 7) { if (yt.k == 4) DW(7, 4) else DW(7, 1) } else if (st.w == 15) DW(15, 1) else if (st.w == 21) DW(21, 1) else DW(27, 1)
#undef DW
#undef DS
  KCHECK(ctx);
  if (side) UG_TRY(stx_join(ctx, xs));
  else if (xblocks > 0) {
    const XPack X{A->xs_ptr, A->xs_len, A->xs_col, A->xs_val};
    if (op == 0) k_dmatmul_xrows<1, 0><<<xblocks, STX_THREADS, 0, ctx->stream>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    else if (op == 1) k_dmatmul_xrows<1, 1><<<xblocks, STX_THREADS, 0, ctx->stream>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    else k_dmatmul_xrows<1, 2><<<xblocks, STX_THREADS, 0, ctx->stream>>>(X, L->n, A->xrows, A->nx, bit, L->ctl, x, y);
    KCHECK(ctx);
  }
  return 0;
}
