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

#include <cstdlib>

#define TR_THREADS 256

// Short rows (P: 1-2 entries on simplices, at most 8 on hexahedra; R: up to 15 / 27), so a warp's life is a chain of dependent
// loads of a few hundred bytes: slice offset -> entries -> gathered operand.  What is done about it:
//  * fixed-width stencils (sell_layout: every uniformly refined level) need no slice-offset load at all;
//  * the head of the kernel (tr_head) is branch-free: all direct-indexed loads of a row are unconditional (indices clamped,
//    results masked), so they leave together (restriction 1.13 -> 0.99 ms, interpolation 1.74 -> 1.62 ms at 513^3);
//  * the far slice's lines are touched for L2 when the warp ends (uggpu_internal.h Prefetch).
// Measured and rejected: K > 1 consecutive slices per warp with the K rows' loads batched -- interpolation 1.62 / 1.86 / 2.25 ms,
// restriction 0.99 / 1.55 / 1.92 ms for K = 1 / 2 / 4 (more registers per thread, fewer resident warps, no gain in bytes in
// flight).  K stays a template parameter and an A/B switch (UGGPU_TR_K_INTERP / UGGPU_TR_K_RESTRICT); the default is 1.
#ifndef TR_K_INTERP
#define TR_K_INTERP 1
#endif
#ifndef TR_K_RESTRICT
#define TR_K_RESTRICT 1
#endif

// L2 prefetch of the stencil entries of the slice pf.dist ahead of row r's (uggpu_internal.h), request and touch in one go
__device__ __forceinline__ bool tr_prefetch(const SellView &T, int r, const Prefetch &pf)
{
  const PfState st = pf_begin(T, r, pf);
  if (T.vcode) {                         // one byte per entry: the slice's codes are val_lines / 8 lines
    if (st.sp >= 0) {
      Prefetch q = pf; q.mode &= ~1;
      pf_end<1>(T, st, q);
      const int lane = threadIdx.x & 31;
      if ((pf.mode & 1) && lane * 128 < pf.val_lines * 16) prefetch_l2(T.vcode + st.sp + (size_t)lane * 128);
    }
  } else pf_end<1>(T, st, pf);
  return (pf.mode & 4) && st.sp >= 0;
}


// Branch-free head of a K-slices-per-warp transfer kernel: every load of the K rows is unconditional (indices clamped into
// range, results masked afterwards), so the compiler can issue the K slices' loads back to back instead of one slice after
// the other (the first K > 1 version wrapped each slice's head in `if (live)` around dependent loads and was SLOWER than K = 1).
template <int K>
struct TrHead {
  int r[K], len[K];
  uint32_t skip[K];
  ColIter ci[K];
  const double *wp[K];
  const uint8_t *cp8[K];     // value codes (T.vcode != nullptr)
  int maxl;
};
// weight of entry j of row k: the stored double, or its one-byte code looked up in the matrix' value table (same bits)
template <int K>
__device__ __forceinline__ double tr_weight(const SellView &T, const TrHead<K> &h, int k, int j)
{
  if (T.vcode) return __ldg(T.vtable + __ldg(h.cp8[k] + (size_t)j * 32));
  return __ldg(h.wp[k] + (size_t)j * 32);
}
template <int K>
__device__ __forceinline__ void tr_head(const SellView &T, const uint32_t *__restrict__ skip_rows, int64_t warp, int lane, TrHead<K> &h)
{
  const int nsl = (T.n + 31) >> 5;
  const bool alias = T.col_ptr == T.slice_ptr;       // no column compression on this matrix (kernel-uniform)
  int64_t sp[K], cp[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int s = (int)min(warp * K + k, (int64_t)nsl - 1);
    sp[k] = T.fixed_w ? (int64_t)s * 32 * T.fixed_w : T.slice_ptr[s];
    cp[k] = alias ? sp[k] : T.col_ptr[s];
  }
  h.maxl = 0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    h.r[k] = (int)((warp * K + k) * 32) + lane;
    const bool live = h.r[k] < T.n;
    const int rr = live ? h.r[k] : 0;
    const int l = T.rowlen[rr];
    const uint32_t sk = skip_rows[rr];
    h.len[k] = live ? l : 0;
    h.skip[k] = sk;
    h.wp[k] = T.val + sp[k] + lane;
    h.cp8[k] = T.vcode + sp[k] + lane;
    const bool uni = cp[k] < 0;
    h.ci[k] = ColIter{uni ? T.col + UG_COLTAB(cp[k]) : T.col + cp[k] + lane, uni ? 1 : 32, uni ? rr : 0};
    h.maxl = max(h.maxl, h.len[k]);
  }
}

// StandardRestrict (transgrid.cc:462 -> :117): to[coarse] (zeroed where VNCLASS >= NEWDEF_CLASS) += sum w * (damp*from[fine]),
// suppressed per component by the coarse VECSKIP bits (:161-165,:180-186).
// Optionally fused (FUSE): the first Jacobi correction of the coarse level, tout = sdamp * Diag(Ac)^-1 to (class-masked,
// ugiter.cc:271 + iter.cc:836) and c = 0 (dset, iter.cc:7873).
template <bool FUSE>
__device__ __forceinline__ void restrict_prefetch(const SellView &R, int r, const Prefetch &pf, const uint8_t *vnclass_c, const uint32_t *skip_c, const uint8_t *vclass_c)
{
  if (tr_prefetch(R, r, pf)) {
    const PfState far{0, -1, (r >> 5) + pf.dist};
    pf_rows<1>(vnclass_c, far, pf); pf_rows<4>(skip_c, far, pf);
    if (FUSE) pf_rows<1>(vclass_c, far, pf);
  }
}

template <int BS, bool FUSE, int K>
__global__ void __launch_bounds__(TR_THREADS) k_restrict_k(SellView R, const uint8_t *__restrict__ vnclass_c, const uint32_t *__restrict__ skip_c,
                                                           double *__restrict__ to, const double *__restrict__ from, Damp damp,
                                                           SellView Ac, const uint8_t *__restrict__ vclass_c, double *__restrict__ tout, double *__restrict__ czero,
                                                           Damp sdamp, int *err, Prefetch pf, HaloK hk)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (hk.flag) halo_publish(hk);                        // multi-GPU (HaloK, uggpu_internal.h; K == 1): ghost rows of the fine defect are read, coarse rows pushed
  if (warp * K * 32 >= R.n) return;
  const uint8_t cf = hk.flag ? hk.flag[warp] : (uint8_t)0;
  TrHead<K> h;
  tr_head<K>(R, skip_c, warp, lane, h);
  if (cf) halo_wait(hk);                                // behind the head's loads (stencil data, not ghost data): no extra round trip for the other warps
  int (&r)[K] = h.r; int (&len)[K] = h.len; uint32_t (&skip)[K] = h.skip; ColIter (&ci)[K] = h.ci;
  const int maxl = h.maxl;
  const bool early = false;       // measured (513^3, B200): touching the far lines when the warp ENDS 1.14 ms, when it starts 1.63 ms
  (void)early;
  double tr[K][BS];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int rr = r[k] < R.n ? r[k] : 0;
    const bool zero = vnclass_c[rr] >= 2;
#pragma unroll
    for (int i = 0; i < BS; i++) { const double t0 = to[(size_t)rr * BS + i]; tr[k][i] = zero ? 0.0 : t0; }
  }
#pragma unroll 4
  for (int j = 0; j < maxl; j++) {
    int f[K];
    double w[K], v[K][BS];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const bool on = j < len[k];
      f[k] = on ? col_at(ci[k], j) : 0;
      w[k] = on ? tr_weight<K>(R, h, k, j) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int i = 0; i < BS; i++) v[k][i] = (cf & 1) ? __ldcg(from + (size_t)f[k] * BS + i) : from[(size_t)f[k] * BS + i];      // rows that are done re-read entry 0 (discarded)
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (j < len[k]) {
#pragma unroll
        for (int i = 0; i < BS; i++) {
          if (!(skip[k] & (1u << i))) {
            double s = damp.a[i] * v[k][i];
            tr[k][i] = tr[k][i] + w[k] * s;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (r[k] >= R.n) continue;
#pragma unroll
    for (int i = 0; i < BS; i++) to[(size_t)r[k] * BS + i] = tr[k][i];
    if (!early) restrict_prefetch<FUSE>(R, r[k], pf, vnclass_c, skip_c, vclass_c);
    if ((cf & 2) && hk.peer && (hk.sel & 255) == HALO_PUSH_B) halo_push_row<BS>(hk, r[k], tr[k]);
    if (FUSE) {
      constexpr int BB = BS * BS;
      double sol[BS];
      if (vclass_c[r[k]] < 3) {
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      } else {
        const double *__restrict__ vp = Ac.diag + ((size_t)(r[k] >> 5) * BB) * 32 + lane;
        double m[BB];
#pragma unroll
        for (int q = 0; q < BB; q++) m[q] = vp[(size_t)q * 32];
        const double (&t)[BS] = tr[k];
        if (BS == 1) sol[0] = t[0] / m[0];
        else {
          // same closed forms as solve_small_block (spmv.cu); duplicated here to keep the kernels self-contained
          if (BS == 2) {
            double det = m[0] * m[3 % BB] - m[1 % BB] * m[2 % BB];
            if (det == 0.0) { atomicExch(err, UGGPU_SMALL_DIAG); det = 1.0; }
            det = 1.0 / det;
            sol[0] = (t[0] * m[3 % BB] - t[1 % BS] * m[1 % BB]) * det;
            sol[1 % BS] = (t[1 % BS] * m[0] - t[0] * m[2 % BB]) * det;
          } else {
            double M3div0 = m[3 % BB] / m[0];
            double M6div0 = m[6 % BB] / m[0];
            double aux = (m[7 % BB] - M6div0 * m[1 % BB]) / (m[4 % BB] - M3div0 * m[1 % BB]);
            sol[2 % BS] = (t[2 % BS] - M6div0 * t[0] - aux * (t[1 % BS] - M3div0 * t[0]))
                          / (m[8 % BB] - M6div0 * m[2 % BB] - aux * (m[5 % BB] - M3div0 * m[2 % BB]));
            sol[1 % BS] = (t[1 % BS] - m[3 % BB] / m[0] * t[0] - (m[5 % BB] - M3div0 * m[2 % BB]) * sol[2 % BS])
                          / (m[4 % BB] - M3div0 * m[1 % BB]);
            sol[0] = (t[0] - m[1 % BB] * sol[1 % BS] - m[2 % BB] * sol[2 % BS]) / m[0];
          }
        }
      }
      double tv[BS];
#pragma unroll
      for (int i = 0; i < BS; i++) {
        tv[i] = sol[i] * sdamp.a[i];
        tout[(size_t)r[k] * BS + i] = tv[i];
        czero[(size_t)r[k] * BS + i] = 0.0;
      }
      if ((cf & 2) && hk.peer && (hk.sel & 255) == HALO_PUSH_TOUT) halo_push_row<BS>(hk, r[k], tv);
    }
  }
}

// StandardInterpolateCorrection (transgrid.cc:529 -> :235): to[fine] = sum (w*damp) * from[coarse], components with the fine
// VECSKIP bit set stay 0 (:272-285).
template <int BS, int K>
__global__ void __launch_bounds__(TR_THREADS, (BS == 1 && K >= 4) ? 4 : ((BS == 1 && K == 1) ? 8 : 1)) k_interpolate_k(SellView P, const uint32_t *__restrict__ skip_f, double *__restrict__ to,
                                                              const double *__restrict__ from, Damp damp, Prefetch pf, HaloK hk)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (hk.flag) halo_publish(hk);                        // multi-GPU (HaloK; K == 1): ghost rows of the coarse correction are read, fine rows pushed
  if (warp * K * 32 >= P.n) return;
  const uint8_t cf = hk.flag ? hk.flag[warp] : (uint8_t)0;
  const bool early = P.fixed_w != 0 && (pf.mode & 64);       // kernel-uniform; opt-in (UGGPU_PF_MODE bit 6): measured equal to touching the lines at the end (1.71 vs 1.74 ms)
  if (early) {
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int rk = (int)((warp * K + k) * 32) + lane;
      if (rk < P.n && tr_prefetch(P, rk, pf)) pf_rows<4>(skip_f, PfState{0, -1, (rk >> 5) + pf.dist}, pf);
    }
  }
  TrHead<K> h;
  tr_head<K>(P, skip_f, warp, lane, h);
  if (cf) halo_wait(hk);                                // behind the head's loads (stencil data, not ghost data)
  int (&r)[K] = h.r; int (&len)[K] = h.len; uint32_t (&skip)[K] = h.skip; ColIter (&ci)[K] = h.ci;
  const int maxl = h.maxl;
  double tr[K][BS];
#pragma unroll
  for (int k = 0; k < K; k++)
#pragma unroll
    for (int i = 0; i < BS; i++) tr[k][i] = 0.0;
#pragma unroll 2
  for (int j = 0; j < maxl; j++) {
    int c[K];
    double w[K], v[K][BS];
#pragma unroll
    for (int k = 0; k < K; k++) {
      const bool on = j < len[k];
      c[k] = on ? col_at(ci[k], j) : 0;
      w[k] = on ? tr_weight<K>(P, h, k, j) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int i = 0; i < BS; i++) v[k][i] = (cf & 1) ? __ldcg(from + (size_t)c[k] * BS + i) : from[(size_t)c[k] * BS + i];      // rows that are done re-read entry 0 (discarded)
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (j < len[k]) {
#pragma unroll
        for (int i = 0; i < BS; i++)
          if (!(skip[k] & (1u << i))) tr[k][i] = tr[k][i] + (w[k] * damp.a[i]) * v[k][i];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (r[k] >= P.n) continue;
#pragma unroll
    for (int i = 0; i < BS; i++) to[(size_t)r[k] * BS + i] = tr[k][i];
    if ((cf & 2) && hk.peer) halo_push_row<BS>(hk, r[k], tr[k]);
    if (!early && tr_prefetch(P, r[k], pf)) pf_rows<4>(skip_f, PfState{0, -1, (r[k] >> 5) + pf.dist}, pf);
  }
}

// grid of a K-slices-per-warp kernel over n rows
template <int K> static inline int tr_blocks(int n)
{
  const int64_t warps = (((int64_t)n + 31) / 32 + K - 1) / K;
  return (int)((warps * 32 + TR_THREADS - 1) / TR_THREADS);
}

// IMAT mode, RestrictByMatrix_General transgrid.cc:1150,1225-1236: the coarse rows with VNCLASS >= NEWDEF_CLASS are scaled after the sums
template <int BS>
__global__ void k_scale_newdef(int n, const uint8_t *__restrict__ vnclass, double *__restrict__ v, Damp damp)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || vnclass[r] < 2) return;
#pragma unroll
  for (int i = 0; i < BS; i++) v[(size_t)r * BS + i] = v[(size_t)r * BS + i] * damp.a[i];
}

static bool damp_is_one(const Damp &d, int bs) { for (int i = 0; i < bs; i++) if (d.a[i] != 1.0) return false; return true; }

extern "C" int uggpu_transfer_set_mode(uggpu_ctx *ctx, int level, int mode)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (mode != UGGPU_TRANSFER_STANDARD && mode != UGGPU_TRANSFER_IMAT) return uggpu_fail(UGGPU_ERROR, "unknown transfer mode %d", mode);
  L->transfer_mode = mode;
  return 0;
}

int k_restrict(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp_in, bool fuse, int A, double *tout, double *czero, Damp sdamp,
               const HaloPlan *hp)
{
  Level *F = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!F || !C) return UGGPU_NO_COARSER_GRID;
  if (!F->R.valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "no transfer stencils on level %d (uggpu_transfer_set)", level);
  // IMAT mode: sums without the damping (1.0 * x is exact), the finished rows scaled afterwards
  const bool post = F->transfer_mode == UGGPU_TRANSFER_IMAT && !damp_is_one(damp_in, F->bs);
  if (post && fuse) return uggpu_fail(UGGPU_ERROR, "restrict: fused Jacobi start with a damped IMAT restriction");
  const Damp damp = F->transfer_mode == UGGPU_TRANSFER_IMAT ? mkdamp(nullptr, 0) : damp_in;
  // partitioned fine level: the coarse rows this rank owns gather from fine ghost rows too.  Peer-memory ghost rows: the kernel waits for
  // the neighbours itself and pushes the coarse rows it produces (hk); otherwise the fine defect is exchanged before the launch.
  HaloK hk = halo_none();
  UG_TRY(halo_prepare(ctx, level - 1, level, &F->R, const_cast<double *>(from), hp, &hk));
  if (hk.peer) {
    hk.sel = (fuse && hp->push == tout) ? HALO_PUSH_TOUT : (hp->push == to ? HALO_PUSH_B : HALO_PUSH_NONE);
    if (hk.sel == HALO_PUSH_NONE) return uggpu_fail(UGGPU_ERROR, "restrict: the vector to push is not produced by this call");
  }
  if (hk.flag && getenv("UGGPU_DBG_HALO")) hk.sel |= atoi(getenv("UGGPU_DBG_HALO"));
  if (C->n == 0 && !hk.flag) return 0;
  const bool gather = ctx->comm && F->partitioned && !C->partitioned;   // first completely held (replicated) level
  if (gather && fuse) return uggpu_fail(UGGPU_ERROR, "restrict: fused Jacobi start not possible across the gather level");
  SellView Rv = view(F->R);
  SellView Av = Rv;
  if (fuse) {
    SellMat *M = get_mat(ctx, level - 1, A);
    if (!M) return UGGPU_DESC_MISMATCH;
    Av = view(*M);
  }
 const double nbf = 8.0 * F->bs * F->n, nbc = 8.0 * F->bs * C->n;
  ProfScope ps(ctx, UGGPU_K_RESTRICT, level, F->R.entry_bytes() + 4.0 * (C->n + 1.0) + nbf + nbc + (fuse ? (double)C->n * 8.0 * F->bs * F->bs + 2.0 * nbc : 0.0));
  // scalar rows: TR_K_RESTRICT slices per warp; block rows already carry BS independent gathers per entry and the fused 3x3 solve
#define RS(BSV, KV)                                                                                                                    \
  {                                                                                                                                    \
    const Prefetch pf = make_prefetch(ctx, &F->R, F->bs, KV);                                                                          \
    const int blocks = tr_blocks<KV>(C->n > 0 ? C->n : 1);                                                                             \
    if (fuse) k_restrict_k<BSV, true, KV><<<blocks, TR_THREADS, 0, ctx->stream>>>(Rv, C->vnclass, C->skip, to, from, damp, Av, C->vclass, tout, czero, sdamp, ctx->derr, pf, hk); \
    else k_restrict_k<BSV, false, KV><<<blocks, TR_THREADS, 0, ctx->stream>>>(Rv, C->vnclass, C->skip, to, from, damp, Av, C->vclass, tout, czero, sdamp, ctx->derr, pf, hk); \
  }
  static const int kenv0 = getenv("UGGPU_TR_K_RESTRICT") ? atoi(getenv("UGGPU_TR_K_RESTRICT")) : TR_K_RESTRICT;     // A/B switch
  const int kenv = hk.flag ? 1 : kenv0;                  // the comm-aware path is the one-slice-per-warp kernel
  int by_class = 0;                                      // rows by class (trc.cu) where the stencil's rows repeat a few shapes
  UG_TRY(trc_restrict(ctx, F, C, to, from, damp, fuse, fuse ? get_mat(ctx, level - 1, A) : nullptr, tout, czero, sdamp, hk, &by_class));
  if (!by_class) {
    switch (F->bs) {
      case 1: if (kenv >= 4) RS(1, 4) else if (kenv >= 2) RS(1, 2) else RS(1, 1) break;
      case 2: RS(2, 1); break;
      default: RS(3, 1); break;
    }
    KCHECK(ctx);
  }
#undef RS
  // every rank filled only the rows of the coarse nodes it would own (the others are 0): summing the disjoint parts
  // is the gather of the coarse defect onto every rank (agglomeration)
  if (gather) UG_TRY(allreduce_sum(ctx, to, (size_t)C->n * C->bs));
  if (post) {
    const int sb = (C->n + 255) / 256;
    switch (F->bs) {
      case 1: k_scale_newdef<1><<<sb, 256, 0, ctx->stream>>>(C->n, C->vnclass, to, damp_in); break;
      case 2: k_scale_newdef<2><<<sb, 256, 0, ctx->stream>>>(C->n, C->vnclass, to, damp_in); break;
      default: k_scale_newdef<3><<<sb, 256, 0, ctx->stream>>>(C->n, C->vnclass, to, damp_in); break;
    }
    KCHECK(ctx);
  }
  return 0;
}

int k_interpolate(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp_in, const HaloPlan *hp)
{
  Level *F = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!F || !C) return UGGPU_NO_COARSER_GRID;
  if (!F->P.valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "no transfer stencils on level %d (uggpu_transfer_set)", level);
  // IMAT mode (InterpolateCorrectionByMatrix_General transgrid.cc:1331,1385): sums without the damping, then dscalx on all rows
  const bool post = F->transfer_mode == UGGPU_TRANSFER_IMAT && !damp_is_one(damp_in, F->bs);
  const Damp damp = F->transfer_mode == UGGPU_TRANSFER_IMAT ? mkdamp(nullptr, 0) : damp_in;
  // coarse ghost values (nothing to do if the coarse level is replicated); peer-memory ghost rows: see k_restrict
  HaloK hk = halo_none();
  UG_TRY(halo_prepare(ctx, level, level - 1, &F->P, const_cast<double *>(from), hp, &hk));
  if (hk.peer && hp->push != to) return uggpu_fail(UGGPU_ERROR, "interpolate: the vector to push is not produced by this call");
  if (hk.flag && getenv("UGGPU_DBG_HALO")) hk.sel |= atoi(getenv("UGGPU_DBG_HALO"));
  if (F->n == 0 && !hk.flag) return 0;
  ProfScope ps(ctx, UGGPU_K_INTERPOLATE, level, F->P.entry_bytes() + 4.0 * (F->n + 1.0) + 8.0 * F->bs * ((double)F->n + C->n));
#define IP(BSV, KV) k_interpolate_k<BSV, KV><<<tr_blocks<KV>(F->n > 0 ? F->n : 1), TR_THREADS, 0, ctx->stream>>>(view(F->P), F->skip, to, from, damp, make_prefetch(ctx, &F->P, F->bs, KV), hk)
  static const int kenv0 = getenv("UGGPU_TR_K_INTERP") ? atoi(getenv("UGGPU_TR_K_INTERP")) : TR_K_INTERP;     // A/B switch
  const int kenv = hk.flag ? 1 : kenv0;
  int by_class = 0;                                      // rows by class (trc.cu)
  UG_TRY(trc_interpolate(ctx, F, C, to, from, damp, hk, &by_class));
  if (!by_class) {
    switch (F->bs) {
      case 1: if (kenv >= 4) IP(1, 4); else if (kenv >= 2) IP(1, 2); else IP(1, 1); break;
      case 2: if (kenv >= 2) IP(2, 2); else IP(2, 1); break;
      default: if (kenv >= 2) IP(3, 2); else IP(3, 1); break;
    }
    KCHECK(ctx);
  }
#undef IP
  if (post) UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, to, nullptr, damp_in));
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
