// blas1.cu -- BLAS level 1 of np/algebra/ugblas.cc:2291-3251 on dense device vectors.
//
// Elementwise ops perform exactly the reference's arithmetic per entry (vecloop.ct:40-49 loop bodies,
// compiled without FMA contraction), so their results are bit-identical.  Reductions use a fixed launch
// geometry and a fixed-order two-stage tree (warp shuffle -> block -> one final block), hence are
// deterministic run to run but not bit-identical to the reference's single running sum.
#include "uggpu_internal.h"

#include <cmath>

// ---------------------------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ double vop(double x, double y, double a)
{
  if (OP == VOP_SET) return a;
  if (OP == VOP_COPY) return y;
  if (OP == VOP_SCALX) return x * a;
  if (OP == VOP_ADD) return x + y;
  if (OP == VOP_SUB) return x - y;
  if (OP == VOP_MINUSADD) return y - x;
  return x + a * y;   // VOP_AXPYX (no contraction: library is compiled with -fmad=false)
}

template <int OP> struct VopTraits { static const bool reads_x = (OP != VOP_SET && OP != VOP_COPY); static const bool reads_y = (OP != VOP_SET && OP != VOP_SCALX); };

// all rows: two entries per thread with 128-bit accesses
template <int OP>
__global__ void __launch_bounds__(256) k_vop_all(size_t cnt, int bs, double *__restrict__ x, const double *__restrict__ y, Damp a)
{
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t k = 2 * p;
  if (k + 1 < cnt) {
    double2 xv = make_double2(0.0, 0.0), yv = make_double2(0.0, 0.0);
    if (VopTraits<OP>::reads_x) xv = *reinterpret_cast<const double2 *>(x + k);
    if (VopTraits<OP>::reads_y) yv = *reinterpret_cast<const double2 *>(y + k);
    int c0 = (int)(k % (size_t)bs), c1 = c0 + 1 == bs ? 0 : c0 + 1;
    double2 r;
    r.x = vop<OP>(xv.x, yv.x, a.a[c0]);
    r.y = vop<OP>(xv.y, yv.y, a.a[c1]);
    *reinterpret_cast<double2 *>(x + k) = r;
  } else if (k < cnt) {
    double xv = VopTraits<OP>::reads_x ? x[k] : 0.0, yv = VopTraits<OP>::reads_y ? y[k] : 0.0;
    x[k] = vop<OP>(xv, yv, a.a[(int)(k % (size_t)bs)]);
  }
}

// masked rows (ON_SURFACE loops, vecloop.ct:22-39): one row per thread
template <int OP>
__global__ void __launch_bounds__(256) k_vop_masked(int n, int bs, uint8_t bit, const uint8_t *__restrict__ ctl, double *__restrict__ x, const double *__restrict__ y, Damp a)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (!(ctl[r] & bit)) return;
  for (int i = 0; i < bs; i++) {
    size_t k = (size_t)r * bs + i;
    double xv = VopTraits<OP>::reads_x ? x[k] : 0.0, yv = VopTraits<OP>::reads_y ? y[k] : 0.0;
    x[k] = vop<OP>(xv, yv, a.a[i]);
  }
}

template <int OP>
static int launch_vop(uggpu_ctx *ctx, Level *L, int rowmode, double *x, const double *y, Damp a)
{
  size_t cnt = (size_t)L->n * L->bs;
  if (cnt == 0) return 0;
  ProfScope ps(ctx, UGGPU_K_VECOP, (int)(L - ctx->lev), 8.0 * cnt * (1.0 + (VopTraits<OP>::reads_x ? 1.0 : 0.0) + (VopTraits<OP>::reads_y ? 1.0 : 0.0)));
  if (rowmode == 0) {
    size_t pairs = (cnt + 1) / 2;
    k_vop_all<OP><<<(unsigned)((pairs + 255) / 256), 256, 0, ctx->stream>>>(cnt, L->bs, x, y, a);
  } else {
    uint8_t bit = rowmode == 1 ? UGGPU_CTL_NEW_DEFECT : UGGPU_CTL_FINE_GRID_DOF;
    k_vop_masked<OP><<<(L->n + 255) / 256, 256, 0, ctx->stream>>>(L->n, L->bs, bit, L->ctl, x, y, a);
  }
  KCHECK(ctx);
  return 0;
}

int k_vec_op(uggpu_ctx *ctx, int level, int rowmode, int op, double *x, const double *y, Damp a)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  switch (op) {
    case VOP_SET: return launch_vop<VOP_SET>(ctx, L, rowmode, x, y, a);
    case VOP_COPY: return launch_vop<VOP_COPY>(ctx, L, rowmode, x, y, a);
    case VOP_SCALX: return launch_vop<VOP_SCALX>(ctx, L, rowmode, x, y, a);
    case VOP_ADD: return launch_vop<VOP_ADD>(ctx, L, rowmode, x, y, a);
    case VOP_SUB: return launch_vop<VOP_SUB>(ctx, L, rowmode, x, y, a);
    case VOP_MINUSADD: return launch_vop<VOP_MINUSADD>(ctx, L, rowmode, x, y, a);
    case VOP_AXPYX: return launch_vop<VOP_AXPYX>(ctx, L, rowmode, x, y, a);
  }
  return uggpu_fail(UGGPU_ERROR, "unknown vector op %d", op);
}

// ---- reductions ---------------------------------------------------------------------------------------------
#define RED_THREADS 256

template <int BS>
__device__ __forceinline__ void block_reduce_store(double (&acc)[BS], double *__restrict__ out /* [BS] for this block */)
{
  __shared__ double sm[RED_THREADS / 32][UGGPU_MAX_BS];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < BS; i++) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[w][i] = v;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double v = lane < RED_THREADS / 32 ? sm[lane][i] : 0.0;
      for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) out[i] = v;
    }
  }
}

template <int BS, int KIND>
__global__ void __launch_bounds__(RED_THREADS) k_red_rows(int n, uint8_t bit, const uint8_t *__restrict__ ctl, const double *__restrict__ x,
                                                          const double *__restrict__ y, double *__restrict__ partials)
{
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    if (bit && !(ctl[r] & bit)) continue;
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double xv = x[(size_t)r * BS + i];
      double yv = KIND == RED_DOT ? y[(size_t)r * BS + i] : xv;
      acc[i] += xv * yv;
    }
  }
  block_reduce_store<BS>(acc, partials + (size_t)blockIdx.x * BS);
}

// final stage: one block sums `count` partial vectors of BS doubles in a fixed order
template <int BS>
__global__ void __launch_bounds__(RED_THREADS) k_red_final(size_t count, const double *__restrict__ partials, double *__restrict__ out)
{
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  for (size_t p = threadIdx.x; p < count; p += RED_THREADS)
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] += partials[p * BS + i];
  block_reduce_store<BS>(acc, out);
}

// middle stage for long partial lists (one partial per block of a row kernel: 10^6 at 513^3): RED_MID blocks sum contiguous
// chunks, in a fixed order, into RED_MID partials behind the list
#define RED_MID 128
template <int BS>
__global__ void __launch_bounds__(RED_THREADS) k_red_mid(size_t count, const double *__restrict__ partials, double *__restrict__ out)
{
  const size_t chunk = (count + RED_MID - 1) / RED_MID;
  const size_t lo = (size_t)blockIdx.x * chunk, hi = lo + chunk < count ? lo + chunk : count;
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  for (size_t p = lo + threadIdx.x; p < hi; p += RED_THREADS)
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] += partials[p * BS + i];
  block_reduce_store<BS>(acc, out + (size_t)blockIdx.x * BS);
}

int reduce_partials_final(uggpu_ctx *ctx, int bs, size_t count, int slot, int level)
{
  double *out = ctx->dres + (size_t)slot * UGGPU_MAX_BS;
  if (count > 16384 && ctx->partials_cap >= (count + RED_MID) * (size_t)bs) {
    double *mid = ctx->partials + count * bs;
    switch (bs) {
      case 1: k_red_mid<1><<<RED_MID, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, mid); k_red_final<1><<<1, RED_THREADS, 0, ctx->stream>>>(RED_MID, mid, out); break;
      case 2: k_red_mid<2><<<RED_MID, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, mid); k_red_final<2><<<1, RED_THREADS, 0, ctx->stream>>>(RED_MID, mid, out); break;
      default: k_red_mid<3><<<RED_MID, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, mid); k_red_final<3><<<1, RED_THREADS, 0, ctx->stream>>>(RED_MID, mid, out); break;
    }
    ctx->launches++;
    KCHECK(ctx);
    if (ctx->comm && level >= 0 && ctx->lev[level].partitioned) UG_TRY(allreduce_sum(ctx, out, (size_t)bs));
    return 0;
  }
  switch (bs) {
    case 1: k_red_final<1><<<1, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, out); break;
    case 2: k_red_final<2><<<1, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, out); break;
    default: k_red_final<3><<<1, RED_THREADS, 0, ctx->stream>>>(count, ctx->partials, out); break;
  }
  KCHECK(ctx);
  // global sum over the ranks (UG_GlobalSumNDOUBLE, parallel/dddif/support.cc:526) where the rows are partitioned
  if (ctx->comm && level >= 0 && ctx->lev[level].partitioned) UG_TRY(allreduce_sum(ctx, out, (size_t)bs));
  return 0;
}

template <int BS>
static int launch_red(uggpu_ctx *ctx, Level *L, int rowmode, int kind, const double *x, const double *y, int slot)
{
  uint8_t bit = rowmode == 0 ? 0 : (rowmode == 1 ? UGGPU_CTL_NEW_DEFECT : UGGPU_CTL_FINE_GRID_DOF);
  int blocks = (L->n + RED_THREADS - 1) / RED_THREADS;
  int cap = ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  UG_TRY(ensure_partials(ctx, (size_t)blocks * BS));
  ProfScope ps(ctx, UGGPU_K_REDUCE, (int)(L - ctx->lev), 8.0 * BS * L->n * (kind == RED_DOT ? 2.0 : 1.0));
  if (kind == RED_DOT) k_red_rows<BS, RED_DOT><<<blocks, RED_THREADS, 0, ctx->stream>>>(L->n, bit, L->ctl, x, y, ctx->partials);
  else k_red_rows<BS, RED_NRM2><<<blocks, RED_THREADS, 0, ctx->stream>>>(L->n, bit, L->ctl, x, y, ctx->partials);
  KCHECK(ctx);
  return reduce_partials_final(ctx, BS, (size_t)blocks, slot, (int)(L - ctx->lev));
}

int k_reduce(uggpu_ctx *ctx, int level, int rowmode, int kind, const double *x, const double *y, int slot)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  switch (L->bs) {
    case 1: return launch_red<1>(ctx, L, rowmode, kind, x, y, slot);
    case 2: return launch_red<2>(ctx, L, rowmode, kind, x, y, slot);
    default: return launch_red<3>(ctx, L, rowmode, kind, x, y, slot);
  }
}

int fetch_results(uggpu_ctx *ctx, int nslots)
{
  CUDA_TRY(cudaMemcpyAsync(ctx->hres, ctx->dres, (size_t)nslots * UGGPU_MAX_BS * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- C-ABI wrappers: loops over levels like vecfunc.ct:23-82 --------------------------------------------------
// (level, rowmode) pairs of one reference loop: ALL_VECTORS = every row of levels fl..tl (vecloop.ct:40-49);
// ON_SURFACE = FINE_GRID_DOF rows of levels FULLREFINELEVEL..tl-1 and NEW_DEFECT rows of tl (vecloop.ct:22-39)
int surface_loop(uggpu_ctx *ctx, int fl, int tl, int mode, std::vector<LoopItem> &out)
{
  out.clear();
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  if (mode == UGGPU_ALL_VECTORS) {
    for (int l = fl; l <= tl; l++) out.push_back({l, 0});
  } else if (mode == UGGPU_ON_SURFACE) {
    for (int l = ctx->fullrefinelevel; l < tl; l++) out.push_back({l, 2});
    out.push_back({tl, 1});
  } else return uggpu_fail(UGGPU_ERROR, "unknown loop mode %d", mode);
  for (auto &it : out) if (!get_level(ctx, it.level)) return UGGPU_ERROR;
  return 0;
}

static int vec_loop(uggpu_ctx *ctx, int fl, int tl, int mode, int op, int x, int y, const double *a, bool scalar_a)
{
  std::vector<LoopItem> items;
  UG_TRY(surface_loop(ctx, fl, tl, mode, items));
  for (auto &it : items) {
    Level *L = &ctx->lev[it.level];
    double *xp = get_vec(ctx, it.level, x);
    if (!xp) return UGGPU_DESC_MISMATCH;
    const double *yp = nullptr;
    if (y >= 0) { yp = get_vec(ctx, it.level, y); if (!yp) return UGGPU_DESC_MISMATCH; }
    Damp d;
    for (int i = 0; i < UGGPU_MAX_BS; i++) d.a[i] = a ? (scalar_a ? a[0] : (i < L->bs ? a[i] : 0.0)) : 0.0;
    UG_TRY(k_vec_op(ctx, it.level, it.rowmode, op, xp, yp, d));
  }
  return 0;
}

extern "C" int uggpu_dset(uggpu_ctx *c, int fl, int tl, int mode, int x, double a) { return vec_loop(c, fl, tl, mode, VOP_SET, x, -1, &a, true); }
extern "C" int uggpu_dcopy(uggpu_ctx *c, int fl, int tl, int mode, int x, int y) { return vec_loop(c, fl, tl, mode, VOP_COPY, x, y, nullptr, true); }
extern "C" int uggpu_dscal(uggpu_ctx *c, int fl, int tl, int mode, int x, double a) { return vec_loop(c, fl, tl, mode, VOP_SCALX, x, -1, &a, true); }
extern "C" int uggpu_dscalx(uggpu_ctx *c, int fl, int tl, int mode, int x, const double *a) { return vec_loop(c, fl, tl, mode, VOP_SCALX, x, -1, a, false); }
extern "C" int uggpu_dadd(uggpu_ctx *c, int fl, int tl, int mode, int x, int y) { return vec_loop(c, fl, tl, mode, VOP_ADD, x, y, nullptr, true); }
extern "C" int uggpu_dsub(uggpu_ctx *c, int fl, int tl, int mode, int x, int y) { return vec_loop(c, fl, tl, mode, VOP_SUB, x, y, nullptr, true); }
extern "C" int uggpu_dminusadd(uggpu_ctx *c, int fl, int tl, int mode, int x, int y) { return vec_loop(c, fl, tl, mode, VOP_MINUSADD, x, y, nullptr, true); }
extern "C" int uggpu_daxpy(uggpu_ctx *c, int fl, int tl, int mode, int x, double a, int y) { return vec_loop(c, fl, tl, mode, VOP_AXPYX, x, y, &a, true); }
extern "C" int uggpu_daxpyx(uggpu_ctx *c, int fl, int tl, int mode, int x, const double *a, int y) { return vec_loop(c, fl, tl, mode, VOP_AXPYX, x, y, a, false); }

// ---- fused chains (uggpu_internal.h CH_*) -----------------------------------------------------------------------------------------------
struct ChainArgs { double *v[5]; double a0, a1; };
template <int CH> struct ChainTraits { static const bool reduces = (CH == CH_ADD_DOT || CH == CH_AXPY2_NRM || CH == CH_BCGS_S); static const int nvec = CH == CH_SCAL_ADD ? 2 : (CH == CH_ADD_DOT ? 3 : (CH == CH_AXPY2_NRM ? 4 : 5)); };

template <int BS, int CH>
__global__ void __launch_bounds__(RED_THREADS) k_chain(int n, int red, uint8_t bit, const uint8_t *__restrict__ ctl, ChainArgs A, double *__restrict__ partials)
{
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const bool in = red && (!bit || (ctl[r] & bit));
#pragma unroll
    for (int i = 0; i < BS; i++) {
      const size_t k = (size_t)r * BS + i;
      if (CH == CH_ADD_DOT) {
        const double b = A.v[0][k] + A.v[1][k];
        A.v[0][k] = b;
        if (in) acc[i] += A.v[2][k] * b;
      } else if (CH == CH_SCAL_ADD) {
        double p = A.v[0][k] * A.a0;
        p = p + A.v[1][k];
        A.v[0][k] = p;
      } else if (CH == CH_AXPY2_NRM) {
        A.v[0][k] = A.v[0][k] + A.a0 * A.v[1][k];
        const double b = A.v[2][k] + A.a1 * A.v[3][k];
        A.v[2][k] = b;
        if (in) acc[i] += b * b;
      } else if (CH == CH_BCGS_P) {
        double p = A.v[0][k] * A.a0;
        p = p + A.v[1][k];
        p = p + A.a1 * A.v[2][k];
        A.v[0][k] = p; A.v[3][k] = 0.0; A.v[4][k] = p;
      } else {   // CH_BCGS_S
        A.v[0][k] = A.v[0][k] + A.a0 * A.v[1][k];
        const double sv = A.v[3][k] + A.a1 * A.v[4][k];
        A.v[2][k] = sv;
        if (in) acc[i] += sv * sv;
      }
    }
  }
  if (ChainTraits<CH>::reduces && red) block_reduce_store<BS>(acc, partials + (size_t)blockIdx.x * BS);
}

template <int BS, int CH>
static int launch_chain(uggpu_ctx *ctx, Level *L, int rowmode, const ChainArgs &A, int slot)
{
  if (L->n == 0) return 0;
  const bool red = ChainTraits<CH>::reduces && rowmode >= 0;
  uint8_t bit = rowmode <= 0 ? 0 : (rowmode == 1 ? UGGPU_CTL_NEW_DEFECT : UGGPU_CTL_FINE_GRID_DOF);
  int blocks = (L->n + RED_THREADS - 1) / RED_THREADS;          // the geometry of launch_red: identical partial sums
  int cap = ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (red) UG_TRY(ensure_partials(ctx, (size_t)blocks * BS));
  {
    ProfScope ps(ctx, UGGPU_K_VECOP, (int)(L - ctx->lev), 8.0 * BS * L->n * (CH == CH_ADD_DOT ? 4.0 : CH == CH_SCAL_ADD ? 3.0 : CH == CH_AXPY2_NRM ? 6.0 : CH == CH_BCGS_P ? 6.0 : 6.0));
    k_chain<BS, CH><<<blocks, RED_THREADS, 0, ctx->stream>>>(L->n, red ? 1 : 0, bit, L->ctl, A, ctx->partials);
    KCHECK(ctx);
  }
  if (red) return reduce_partials_final(ctx, BS, (size_t)blocks, slot, (int)(L - ctx->lev));
  return 0;
}

template <int CH>
static int chain_level(uggpu_ctx *ctx, Level *L, int rowmode, const ChainArgs &A, int slot)
{
  switch (L->bs) {
    case 1: return launch_chain<1, CH>(ctx, L, rowmode, A, slot);
    case 2: return launch_chain<2, CH>(ctx, L, rowmode, A, slot);
    default: return launch_chain<3, CH>(ctx, L, rowmode, A, slot);
  }
}

int chain_loop(uggpu_ctx *ctx, int fl, int tl, int chain, const int *vecs, double a0, double a1, double *sums, int *bs_out)
{
  std::vector<LoopItem> surf;
  UG_TRY(surface_loop(ctx, fl, tl, UGGPU_ON_SURFACE, surf));
  const int nvec = chain == CH_SCAL_ADD ? 2 : (chain == CH_ADD_DOT ? 3 : (chain == CH_AXPY2_NRM ? 4 : 5));
  int bs = 1, nslots = 0;
  // the reduction's slots follow reduce_loop: one per surface level, ascending
  int slot_of[UGGPU_MAX_LEVELS], mode_of[UGGPU_MAX_LEVELS];
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) { slot_of[l] = -1; mode_of[l] = -1; }
  for (auto &it : surf) { slot_of[it.level] = nslots++; mode_of[it.level] = it.rowmode; }
  for (int l = fl; l <= tl; l++) {
    Level *L = get_level(ctx, l);
    if (!L) return UGGPU_ERROR;
    bs = L->bs;
    ChainArgs A;
    for (int i = 0; i < 5; i++) A.v[i] = nullptr;
    for (int i = 0; i < nvec; i++) { A.v[i] = get_vec(ctx, l, vecs[i]); if (!A.v[i]) return UGGPU_DESC_MISMATCH; }
    A.a0 = a0; A.a1 = a1;
    const int rm = mode_of[l], sl = slot_of[l] < 0 ? 0 : slot_of[l];
    switch (chain) {
      case CH_ADD_DOT: UG_TRY(chain_level<CH_ADD_DOT>(ctx, L, rm, A, sl)); break;
      case CH_SCAL_ADD: UG_TRY(chain_level<CH_SCAL_ADD>(ctx, L, rm, A, sl)); break;
      case CH_AXPY2_NRM: UG_TRY(chain_level<CH_AXPY2_NRM>(ctx, L, rm, A, sl)); break;
      case CH_BCGS_P: UG_TRY(chain_level<CH_BCGS_P>(ctx, L, rm, A, sl)); break;
      case CH_BCGS_S: UG_TRY(chain_level<CH_BCGS_S>(ctx, L, rm, A, sl)); break;
      default: return uggpu_fail(UGGPU_ERROR, "unknown chain %d", chain);
    }
  }
  if (bs_out) *bs_out = bs;
  if (!sums || chain == CH_SCAL_ADD || chain == CH_BCGS_P) return 0;
  // surface levels below fl (ON_SURFACE starts at FULLREFINELEVEL whatever fl is, vecloop.ct:22): their rows are not touched by the
  // chain's vector operations but belong to the reduction -- callers use fl <= FULLREFINELEVEL, anything else is refused
  for (auto &it : surf) if (it.level < fl) return uggpu_fail(UGGPU_ERROR, "chain_loop: surface level %d below the first level %d", it.level, fl);
  UG_TRY(fetch_results(ctx, nslots));
  for (int i = 0; i < UGGPU_MAX_BS; i++) sums[i] = 0.0;
  for (int s2 = 0; s2 < nslots; s2++)
    for (int i = 0; i < bs; i++) sums[i] += ctx->hres[s2 * UGGPU_MAX_BS + i];
  return 0;
}

// sums[i] = sum over the loop of x_i*y_i per component i (ddotx ugblas.cc:2946) -- host adds the per-level results in level order
int reduce_loop(uggpu_ctx *ctx, int fl, int tl, int mode, int kind, int x, int y, double *sums /* [MAX_BS] */, int *bs_out)
{
  std::vector<LoopItem> items;
  UG_TRY(surface_loop(ctx, fl, tl, mode, items));
  if ((int)items.size() > UGGPU_MAX_LEVELS) return uggpu_fail(UGGPU_ERROR, "too many levels");
  int slot = 0, bs = 1;
  for (auto &it : items) {
    const double *xp = get_vec(ctx, it.level, x);
    const double *yp = kind == RED_DOT ? get_vec(ctx, it.level, y) : xp;
    if (!xp || !yp) return UGGPU_DESC_MISMATCH;
    bs = ctx->lev[it.level].bs;
    UG_TRY(k_reduce(ctx, it.level, it.rowmode, kind, xp, yp, slot++));
  }
  UG_TRY(fetch_results(ctx, slot));
  for (int i = 0; i < UGGPU_MAX_BS; i++) sums[i] = 0.0;
  for (int s = 0; s < slot; s++)
    for (int i = 0; i < bs; i++) sums[i] += ctx->hres[s * UGGPU_MAX_BS + i];
  if (bs_out) *bs_out = bs;
  return 0;
}

extern "C" int uggpu_ddotx(uggpu_ctx *c, int fl, int tl, int mode, int x, int y, double *a)
{
  double s[UGGPU_MAX_BS]; int bs;
  UG_TRY(reduce_loop(c, fl, tl, mode, RED_DOT, x, y, s, &bs));
  for (int i = 0; i < bs; i++) a[i] = s[i];
  return 0;
}
extern "C" int uggpu_ddot(uggpu_ctx *c, int fl, int tl, int mode, int x, int y, double *a)
{
  double s[UGGPU_MAX_BS]; int bs;
  UG_TRY(reduce_loop(c, fl, tl, mode, RED_DOT, x, y, s, &bs));
  double t = 0.0;
  for (int i = 0; i < bs; i++) t += s[i];
  *a = t;
  return 0;
}
extern "C" int uggpu_dnrm2x(uggpu_ctx *c, int fl, int tl, int mode, int x, double *a)
{
  double s[UGGPU_MAX_BS]; int bs;
  UG_TRY(reduce_loop(c, fl, tl, mode, RED_NRM2, x, x, s, &bs));
  for (int i = 0; i < bs; i++) a[i] = sqrt(s[i]);      // SQRT after the sum, ugblas.cc:3171
  return 0;
}
extern "C" int uggpu_dnrm2(uggpu_ctx *c, int fl, int tl, int mode, int x, double *a)
{
  double s[UGGPU_MAX_BS]; int bs;
  UG_TRY(reduce_loop(c, fl, tl, mode, RED_NRM2, x, x, s, &bs));
  double t = 0.0;
  for (int i = 0; i < bs; i++) t += s[i];
  *a = sqrt(t);
  return 0;
}
