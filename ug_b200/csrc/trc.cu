// trc.cu -- grid transfer by ROW CLASS: StandardRestrict / StandardInterpolateCorrection (np/algebra/transgrid.cc:117-336) on
// transfer stencils whose rows fall into a few classes.
//
// The stored stencils P (rows = fine vectors) and R (rows = coarse vectors) are general sparse matrices (transfer.cu works on them
// entry by entry: 4-byte column + 1-byte weight code per entry, row length, per-row flags).  On any hierarchy produced by regular
// refinement most rows repeat one of a handful of shapes: the same number of entries, the same column DISTANCES to the row's first
// entry, the same weights, the same flags -- a fine node is a copy of a coarse node or the midpoint of one of 7 edge directions
// (simplices; 26 + 1 shapes on hexahedra), times the boundary variants.  Here a row is stored as
//     base[r]  (int32)  the column of its first entry           cls[r]  (uint8)  its class, 255 = exception row
// and a class as one record (length, distances, weights, skip bits, flags) in a table of at most 255 records: 5 bytes per row
// instead of 12 per entry + 6 per row, one round trip (base, class) before the gathers instead of a chain through row length,
// column words and weight codes.  Restriction on the finest level of the 513^3 hierarchy: 77 -> 5 bytes of stencil per coarse row.
//
// Lossless and generic: classes are FOUND on the device (hash of the row's shape into a 256-slot table, the first row of a shape is its
// representative), the table is filled from the representatives, and every row is then compared with its class record entry by
// entry and bit by bit -- a row that differs (hash collision), a row of a 256th shape, a row longer than 27 entries becomes an
// EXCEPTION row (class 255).  Exception rows are processed from the general stencil by a second kernel, one thread per row, on a
// second stream.  Multi-GPU (peer-memory ghost rows): rows with ghost columns and rows to push are exception rows as in stx.cu, so the
// class kernel never waits and never stores remotely.  Hierarchies without repeating shapes (adaptive, unstructured) simply keep
// transfer.cu's kernels: more than half of the rows in exception -> not used.
//
// Same entries, same order, same arithmetic as transfer.cu: bit-identical results (tests: UGGPU_NO_TRC=1 A/B, port, golden dumps).
#include "uggpu_internal.h"

#include <cstdlib>
#include <vector>

#define TRC_MAXLEN 27
#define TRC_SLOTS 256
#define TRC_EXC 255
#define TRC_THREADS 256

struct TrClass {
  int len;
  uint32_t skip;             // VECSKIP bits of the row's vector
  uint32_t aux;              // restriction: bit 0 coarse VNCLASS >= NEWDEF_CLASS (accumulator starts at 0), bit 1 coarse VCLASS < ACTIVE_CLASS (fused Jacobi start gives 0)
  int delta[TRC_MAXLEN];     // column of entry j minus column of entry 0
  double w[TRC_MAXLEN];
};

struct TrcData {
  int32_t *base = nullptr;
  uint8_t *cls = nullptr;
  TrClass *table = nullptr;
  int32_t *xrows = nullptr;
  int n = 0, nx = 0, ncls = 0;
  int comm = 0;              // built with the multi-GPU exceptions
  bool usable = false;
};

// entry j of row r of a transfer stencil: column and weight (the stored double, or its one-byte code looked up in the value table)
struct TrRow { int64_t sp; ColIter ci; int lane; };
__device__ __forceinline__ TrRow tr_row(const SellView &T, int r)
{
  return TrRow{slice_off(T, r >> 5), col_iter(T, r), r & 31};
}
__device__ __forceinline__ double tr_w(const SellView &T, const TrRow &q, int j)
{
  const int64_t e = q.sp + (int64_t)j * 32 + q.lane;
  return T.vcode ? __ldg(T.vtable + __ldg(T.vcode + e)) : __ldg(T.val + e);
}

__device__ __forceinline__ unsigned long long trc_mix(unsigned long long h, unsigned long long v)
{
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h * 0xBF58476D1CE4E5B9ull;
}

// shape of row r -> class slot; rows that cannot have a class (too long, multi-GPU exceptions) -> TRC_EXC
__global__ void k_trc_classify(SellView T, const uint32_t *__restrict__ skip_rows, const uint8_t *__restrict__ vnclass, const uint8_t *__restrict__ vclass,
                               int n_owned_cols, const uint32_t *__restrict__ snd_bits, unsigned long long *slots, int *rep, int *overflow,
                               int32_t *__restrict__ base, uint8_t *__restrict__ cls)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= T.n) return;
  const int len = T.rowlen[r];
  const TrRow q = tr_row(T, r);
  const int c0 = len > 0 ? col_at(q.ci, 0) : 0;
  base[r] = c0;
  bool exc = len > TRC_MAXLEN;
  if (snd_bits && ((snd_bits[r >> 5] >> (r & 31)) & 1u)) exc = true;
  const uint32_t aux = (vnclass && vnclass[r] >= 2 ? 1u : 0u) | (vclass && vclass[r] < 3 ? 2u : 0u);
  unsigned long long h = trc_mix(0x1234567ull, (unsigned long long)len);
  h = trc_mix(h, ((unsigned long long)skip_rows[r] << 8) | aux);
  for (int j = 0; j < len && !exc; j++) {
    const int c = col_at(q.ci, j);
    if (n_owned_cols >= 0 && c >= n_owned_cols) exc = true;
    h = trc_mix(h, (unsigned long long)(unsigned int)(c - c0));
    h = trc_mix(h, (unsigned long long)__double_as_longlong(tr_w(T, q, j)));
  }
  if (exc) { cls[r] = TRC_EXC; return; }
  if (h == 0ull) h = 1ull;
  unsigned s = (unsigned)(h >> 40) % (TRC_SLOTS - 1);
  for (int probe = 0; probe < TRC_SLOTS - 1; probe++) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(slots + s);
    if (cur == 0ull) {
      cur = atomicCAS(slots + s, 0ull, h);
      if (cur == 0ull) { atomicMin(rep + s, r); cls[r] = (uint8_t)s; return; }
    }
    if (cur == h) { atomicMin(rep + s, r); cls[r] = (uint8_t)s; return; }
    s = (s + 1) % (TRC_SLOTS - 1);
  }
  atomicAdd(overflow, 1);            // a 256th shape: exception row
  cls[r] = TRC_EXC;
}

// class record of slot s from its representative row (the lowest row index that landed in the slot)
__global__ void k_trc_fill(SellView T, const uint32_t *__restrict__ skip_rows, const uint8_t *__restrict__ vnclass, const uint8_t *__restrict__ vclass,
                           const unsigned long long *__restrict__ slots, const int *__restrict__ rep, TrClass *__restrict__ table)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= TRC_SLOTS) return;
  TrClass t;
  t.len = 0; t.skip = 0; t.aux = 0;
  for (int j = 0; j < TRC_MAXLEN; j++) { t.delta[j] = 0; t.w[j] = 0.0; }
  if (s < TRC_SLOTS - 1 && slots[s] != 0ull) {
    const int r = rep[s];
    const TrRow q = tr_row(T, r);
    t.len = T.rowlen[r];
    t.skip = skip_rows[r];
    t.aux = (vnclass && vnclass[r] >= 2 ? 1u : 0u) | (vclass && vclass[r] < 3 ? 2u : 0u);
    const int c0 = t.len > 0 ? col_at(q.ci, 0) : 0;
    for (int j = 0; j < t.len; j++) { t.delta[j] = col_at(q.ci, j) - c0; t.w[j] = tr_w(T, q, j); }
  }
  table[s] = t;
}

// every classed row against its class record, entry by entry and bit by bit; a row that differs becomes an exception row.
// xbits: bit l of word s = row 32 s + l is an exception row
__global__ void k_trc_verify(SellView T, const uint32_t *__restrict__ skip_rows, const uint8_t *__restrict__ vnclass, const uint8_t *__restrict__ vclass,
                             const TrClass *__restrict__ table, const int32_t *__restrict__ base, uint8_t *__restrict__ cls, uint32_t *__restrict__ xbits)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  bool exc = false;
  if (r < T.n) {
    const int c = cls[r];
    exc = c == TRC_EXC;
    if (!exc) {
      const TrClass &t = table[c];
      const int len = T.rowlen[r];
      const uint32_t aux = (vnclass && vnclass[r] >= 2 ? 1u : 0u) | (vclass && vclass[r] < 3 ? 2u : 0u);
      bool same = len == t.len && skip_rows[r] == t.skip && aux == t.aux;
      if (same) {
        const TrRow q = tr_row(T, r);
        for (int j = 0; j < len && same; j++)
          if (col_at(q.ci, j) - base[r] != t.delta[j] || __double_as_longlong(tr_w(T, q, j)) != __double_as_longlong(t.w[j])) same = false;
      }
      if (!same) { cls[r] = TRC_EXC; exc = true; }
    }
  }
  const uint32_t m = __ballot_sync(0xffffffffu, exc);
  if ((threadIdx.x & 31) == 0 && (r & ~31) < T.n) xbits[r >> 5] = m;
}

int trc_free(uggpu_ctx *ctx, SellMat *m)
{
  TrcData *d = m->trc;
  if (!d) return 0;
  if (d->base) dfree(ctx, d->base, (size_t)d->n + 1);
  if (d->cls) dfree(ctx, d->cls, (size_t)d->n + 1);
  if (d->table) dfree(ctx, d->table, (size_t)TRC_SLOTS);
  if (d->xrows) dfree(ctx, d->xrows, (size_t)d->nx + 1);
  delete d;
  m->trc = nullptr;
  return 0;
}

int trc_free_comm(uggpu_ctx *ctx, SellMat *m) { return (m->trc && m->trc->comm) ? trc_free(ctx, m) : 0; }

// T: P or R; rows on level RL (flags skip / vnclass / vclass of those rows; the last two only matter for R); n_owned_cols: columns
// beyond are ghost columns (-1: none); snd: rows to push (nullptr: none).  comm: build with the multi-GPU exceptions.
static int trc_ensure(uggpu_ctx *ctx, SellMat *T, const Level *RL, bool is_R, int n_owned_cols, const uint32_t *snd, bool comm, TrcData **out)
{
  *out = nullptr;
  if (T->trc && T->trc->comm == (comm ? 1 : 0)) { if (T->trc->usable) *out = T->trc; return 0; }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  UG_TRY(trc_free(ctx, T));
  TrcData *d = new TrcData();
  T->trc = d;
  d->n = T->n; d->comm = comm ? 1 : 0;
  if (T->n < 1 || T->bb != 1 || T->maxlen < 1) return 0;
  cudaStream_t st = ctx->stream;
  const size_t nsl = ((size_t)T->n + 31) / 32;
  unsigned long long *slots = nullptr; int *rep = nullptr, *ovf = nullptr; uint32_t *xbits = nullptr;
  UG_TRY(dalloc(ctx, &d->base, (size_t)d->n + 1));
  UG_TRY(dalloc(ctx, &d->cls, (size_t)d->n + 1));
  UG_TRY(dalloc(ctx, &d->table, (size_t)TRC_SLOTS));
  UG_TRY(dalloc(ctx, &slots, (size_t)TRC_SLOTS));
  UG_TRY(dalloc(ctx, &rep, (size_t)TRC_SLOTS));
  UG_TRY(dalloc(ctx, &ovf, 1));
  UG_TRY(dalloc(ctx, &xbits, nsl + 1));
  CUDA_TRY(cudaMemsetAsync(slots, 0, sizeof(unsigned long long) * TRC_SLOTS, st));
  CUDA_TRY(cudaMemsetAsync(rep, 0x7f, sizeof(int) * TRC_SLOTS, st));
  CUDA_TRY(cudaMemsetAsync(ovf, 0, sizeof(int), st));
  const int blocks = (T->n + TRC_THREADS - 1) / TRC_THREADS;
  const uint8_t *vn = is_R ? RL->vnclass : nullptr, *vc = is_R ? RL->vclass : nullptr;
  k_trc_classify<<<blocks, TRC_THREADS, 0, st>>>(view(*T), RL->skip, vn, vc, comm ? n_owned_cols : -1, comm ? snd : nullptr, slots, rep, ovf, d->base, d->cls);
  KCHECK(ctx);
  k_trc_fill<<<1, TRC_SLOTS, 0, st>>>(view(*T), RL->skip, vn, vc, slots, rep, d->table);
  KCHECK(ctx);
  k_trc_verify<<<(int)((nsl * 32 + TRC_THREADS - 1) / TRC_THREADS), TRC_THREADS, 0, st>>>(view(*T), RL->skip, vn, vc, d->table, d->base, d->cls, xbits);
  KCHECK(ctx);
  std::vector<uint32_t> bits(nsl);
  std::vector<unsigned long long> hs(TRC_SLOTS);
  CUDA_TRY(cudaMemcpyAsync(bits.data(), xbits, sizeof(uint32_t) * nsl, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(hs.data(), slots, sizeof(unsigned long long) * TRC_SLOTS, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  dfree(ctx, slots, (size_t)TRC_SLOTS); dfree(ctx, rep, (size_t)TRC_SLOTS); dfree(ctx, ovf, 1); dfree(ctx, xbits, nsl + 1);
  for (auto h : hs) if (h) d->ncls++;
  std::vector<int32_t> rows;
  for (size_t s = 0; s < nsl; s++) {
    const uint32_t m = bits[s];
    if (!m) continue;
    for (int l = 0; l < 32; l++) if ((m >> l) & 1u) rows.push_back((int32_t)(s * 32 + l));
  }
  d->nx = (int)rows.size();
  // worth it only when most rows have a class (regular refinement); otherwise transfer.cu's kernels stay
  if ((int64_t)d->nx * 2 > (int64_t)d->n) return 0;
  UG_TRY(dalloc(ctx, &d->xrows, rows.size() + 1));
  if (!rows.empty()) CUDA_TRY(cudaMemcpyAsync(d->xrows, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  d->usable = true;
  *out = d;
  return 0;
}

// ---- class kernels -----------------------------------------------------------------------------------------------------------------
// Interpolation: TWO consecutive rows per thread and the loads in three waves -- (1) base and class of both rows, (2) the class records'
// heads AND the first gathers (entry 0 sits at the base column itself: delta[0] = 0), (3) the second gathers -- instead of one row with a
// load-use stall per entry: the kernel moves 13 bytes per row and lives on how few times a warp has to wait (ncu, one row per thread:
// 30 of 37 stall cycles per instruction on the L1TEX scoreboard, DRAM at 25 %).
template <int BS>
__global__ void __launch_bounds__(TRC_THREADS) k_interp_cls(int n, const int32_t *__restrict__ base, const uint8_t *__restrict__ cls, const TrClass *__restrict__ table,
                                                            double *__restrict__ to, const double *__restrict__ from, Damp damp)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = 2 * t;
  if (r0 >= n) return;
  const bool two = r0 + 1 < n;
  int b[2], c[2];
  if (two) {
    const int2 bb = __ldg(reinterpret_cast<const int2 *>(base) + t);
    const uchar2 cc = __ldg(reinterpret_cast<const uchar2 *>(cls) + t);
    b[0] = bb.x; b[1] = bb.y; c[0] = cc.x; c[1] = cc.y;
  } else { b[0] = base[r0]; c[0] = cls[r0]; b[1] = 0; c[1] = TRC_EXC; }
  int len[2]; uint32_t skip[2]; int d1[2]; double w0[2], w1[2], v0[2][BS], v1[2][BS], tr[2][BS];
#pragma unroll
  for (int q = 0; q < 2; q++) {
    const bool on = c[q] != TRC_EXC;
    const TrClass &tc = table[on ? c[q] : 0];
    len[q] = on ? tc.len : 0; skip[q] = tc.skip; d1[q] = tc.delta[1]; w0[q] = tc.w[0]; w1[q] = tc.w[1];
#pragma unroll
    for (int i = 0; i < BS; i++) v0[q][i] = from[(size_t)(on ? b[q] : 0) * BS + i];      // entry 0: no table needed for its address
  }
#pragma unroll
  for (int q = 0; q < 2; q++)
#pragma unroll
    for (int i = 0; i < BS; i++) v1[q][i] = from[(size_t)(len[q] >= 2 ? b[q] + d1[q] : 0) * BS + i];
#pragma unroll
  for (int q = 0; q < 2; q++) {
#pragma unroll
    for (int i = 0; i < BS; i++) {
      double a = 0.0;
      if (!(skip[q] & (1u << i))) {
        if (len[q] >= 1) a = a + (w0[q] * damp.a[i]) * v0[q][i];
        if (len[q] >= 2) a = a + (w1[q] * damp.a[i]) * v1[q][i];
      }
      tr[q][i] = a;
    }
    if (len[q] > 2) {                     // hexahedra: up to 8 entries
      const TrClass &tc = table[c[q]];
      for (int j = 2; j < len[q]; j++) {
        const int col = b[q] + tc.delta[j];
        const double w = tc.w[j];
#pragma unroll
        for (int i = 0; i < BS; i++) {
          const double v = from[(size_t)col * BS + i];
          if (!(skip[q] & (1u << i))) tr[q][i] = tr[q][i] + (w * damp.a[i]) * v;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 2; q++)
    if (c[q] != TRC_EXC) {
#pragma unroll
      for (int i = 0; i < BS; i++) to[(size_t)(r0 + q) * BS + i] = tr[q][i];
    }
}

template <int BS, bool FUSE>
__global__ void __launch_bounds__(TRC_THREADS) k_restrict_cls(int n, const int32_t *__restrict__ base, const uint8_t *__restrict__ cls, const TrClass *__restrict__ table,
                                                              double *__restrict__ to, const double *__restrict__ from, Damp damp,
                                                              const double *__restrict__ diag, double *__restrict__ tout, double *__restrict__ czero, Damp sdamp, int *err)
{
  constexpr int BB = BS * BS;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int c = cls[r];
  const int b = base[r];
  if (c == TRC_EXC) return;
  const TrClass &t = table[c];
  const int len = t.len;
  const uint32_t skip = t.skip, aux = t.aux;
  double tr[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) { const double t0 = to[(size_t)r * BS + i]; tr[i] = (aux & 1u) ? 0.0 : t0; }
#pragma unroll 4
  for (int j = 0; j < len; j++) {
    const int f = b + t.delta[j];
    const double w = t.w[j];
#pragma unroll
    for (int i = 0; i < BS; i++) {
      const double v = from[(size_t)f * BS + i];
      if (!(skip & (1u << i))) { const double s = damp.a[i] * v; tr[i] = tr[i] + w * s; }
    }
  }
#pragma unroll
  for (int i = 0; i < BS; i++) to[(size_t)r * BS + i] = tr[i];
  if (FUSE) {
    double sol[BS];
    if (aux & 2u) {
#pragma unroll
      for (int i = 0; i < BS; i++) sol[i] = 0.0;
    } else {
      const double *__restrict__ vp = diag + ((size_t)(r >> 5) * BB) * 32 + (r & 31);
      double m[BB];
#pragma unroll
      for (int q = 0; q < BB; q++) m[q] = vp[(size_t)q * 32];
      if (solve_small_block<BS>(m, tr, sol)) {
        atomicExch(err, UGGPU_SMALL_DIAG);
        // transfer.cu's fused form divides by a determinant forced to 1 for a singular 2x2 block; the error word is what counts
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < BS; i++) { tout[(size_t)r * BS + i] = sol[i] * sdamp.a[i]; czero[(size_t)r * BS + i] = 0.0; }
  }
}

// ---- exception rows: one thread per row on the general stencil -----------------------------------------------------------------------
template <int BS, bool COMM>
__global__ void __launch_bounds__(128) k_interp_xrows(SellView P, const int32_t *__restrict__ xrows, int nx, int n_owned_cols, const uint32_t *__restrict__ skip_f,
                                                      double *__restrict__ to, const double *__restrict__ from, Damp damp, HaloK hk)
{
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < nx;
  const int r = live ? __ldg(xrows + i0) : 0;
  if (COMM) { halo_publish(hk); halo_wait(hk); }
  if (!live) return;
  const TrRow q = tr_row(P, r);
  const int len = P.rowlen[r];
  const uint32_t skip = skip_f[r];
  double tr[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) tr[i] = 0.0;
#pragma unroll 2
  for (int j = 0; j < len; j++) {
    const int col = col_at(q.ci, j);
    const double w = tr_w(P, q, j);
#pragma unroll
    for (int i = 0; i < BS; i++) {
      const double v = (COMM && col >= n_owned_cols) ? __ldcg(from + (size_t)col * BS + i) : from[(size_t)col * BS + i];
      if (!(skip & (1u << i))) tr[i] = tr[i] + (w * damp.a[i]) * v;
    }
  }
#pragma unroll
  for (int i = 0; i < BS; i++) to[(size_t)r * BS + i] = tr[i];
  if (COMM && hk.peer) halo_push_row<BS>(hk, r, tr);
}

template <int BS, bool FUSE, bool COMM>
__global__ void __launch_bounds__(128) k_restrict_xrows(SellView R, const int32_t *__restrict__ xrows, int nx, int n_owned_cols, const uint8_t *__restrict__ vnclass_c,
                                                        const uint32_t *__restrict__ skip_c, double *__restrict__ to, const double *__restrict__ from, Damp damp,
                                                        const double *__restrict__ diag, const uint8_t *__restrict__ vclass_c, double *__restrict__ tout,
                                                        double *__restrict__ czero, Damp sdamp, int *err, HaloK hk)
{
  constexpr int BB = BS * BS;
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i0 < nx;
  const int r = live ? __ldg(xrows + i0) : 0;
  if (COMM) { halo_publish(hk); halo_wait(hk); }
  if (!live) return;
  const TrRow q = tr_row(R, r);
  const int len = R.rowlen[r];
  const uint32_t skip = skip_c[r];
  const bool zero = vnclass_c[r] >= 2;
  double tr[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) { const double t0 = to[(size_t)r * BS + i]; tr[i] = zero ? 0.0 : t0; }
#pragma unroll 4
  for (int j = 0; j < len; j++) {
    const int f = col_at(q.ci, j);
    const double w = tr_w(R, q, j);
#pragma unroll
    for (int i = 0; i < BS; i++) {
      const double v = (COMM && f >= n_owned_cols) ? __ldcg(from + (size_t)f * BS + i) : from[(size_t)f * BS + i];
      if (!(skip & (1u << i))) { const double s = damp.a[i] * v; tr[i] = tr[i] + w * s; }
    }
  }
#pragma unroll
  for (int i = 0; i < BS; i++) to[(size_t)r * BS + i] = tr[i];
  if (COMM && hk.peer && (hk.sel & 255) == HALO_PUSH_B) halo_push_row<BS>(hk, r, tr);
  if (FUSE) {
    double sol[BS];
    if (vclass_c[r] < 3) {
#pragma unroll
      for (int i = 0; i < BS; i++) sol[i] = 0.0;
    } else {
      const double *__restrict__ vp = diag + ((size_t)(r >> 5) * BB) * 32 + (r & 31);
      double m[BB];
#pragma unroll
      for (int k = 0; k < BB; k++) m[k] = vp[(size_t)k * 32];
      if (solve_small_block<BS>(m, tr, sol)) {
        atomicExch(err, UGGPU_SMALL_DIAG);
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      }
    }
    double tv[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) { tv[i] = sol[i] * sdamp.a[i]; tout[(size_t)r * BS + i] = tv[i]; czero[(size_t)r * BS + i] = 0.0; }
    if (COMM && hk.peer && (hk.sel & 255) == HALO_PUSH_TOUT) halo_push_row<BS>(hk, r, tv);
  }
}

// ---- launches ---------------------------------------------------------------------------------------------------------------------------
static bool trc_enabled(int n)
{
  if (getenv("UGGPU_NO_TRC")) return false;
  const char *mr = getenv("UGGPU_TRC_MIN_ROWS");
  return n >= (mr ? atoi(mr) : 4096);
}

static int side_begin(uggpu_ctx *ctx, cudaStream_t *xs)
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
static int side_end(uggpu_ctx *ctx, cudaStream_t xs)
{
  CUDA_TRY(cudaEventRecord(ctx->halo_ev[1], xs));
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->halo_ev[1], 0));
  return 0;
}

int trc_interpolate(uggpu_ctx *ctx, Level *F, Level *C, double *to, const double *from, Damp damp, const HaloK &hk, int *done)
{
  *done = 0;
  if (!trc_enabled(F->n)) return 0;
  const bool comm = hk.flag != nullptr;
  const bool cpart = ctx->comm && C->partitioned;
  TrcData *d = nullptr;
  UG_TRY(trc_ensure(ctx, &F->P, F, false, cpart ? C->n : -1, comm ? halo_snd_bits(F) : nullptr, comm, &d));
  if (!d) return 0;
  *done = 1;
  cudaStream_t xs = ctx->stream;
  const bool side = d->nx > 0 || comm;
  if (side) UG_TRY(side_begin(ctx, &xs));
  const int xblocks = (d->nx + 127) / 128 > 0 ? (d->nx + 127) / 128 : 1;
  const int nown = cpart ? C->n : 0x7fffffff;
#define XI(BSV) { if (comm) k_interp_xrows<BSV, true><<<xblocks, 128, 0, xs>>>(view(F->P), d->xrows, d->nx, nown, F->skip, to, from, damp, hk); \
                  else k_interp_xrows<BSV, false><<<xblocks, 128, 0, xs>>>(view(F->P), d->xrows, d->nx, nown, F->skip, to, from, damp, hk); }
  if (side) { switch (F->bs) { case 1: XI(1) break; case 2: XI(2) break; default: XI(3) break; } KCHECK(ctx); }
#undef XI
  const int blocks = ((F->n + 1) / 2 + TRC_THREADS - 1) / TRC_THREADS;      // two rows per thread
  switch (F->bs) {
    case 1: k_interp_cls<1><<<blocks, TRC_THREADS, 0, ctx->stream>>>(F->n, d->base, d->cls, d->table, to, from, damp); break;
    case 2: k_interp_cls<2><<<blocks, TRC_THREADS, 0, ctx->stream>>>(F->n, d->base, d->cls, d->table, to, from, damp); break;
    default: k_interp_cls<3><<<blocks, TRC_THREADS, 0, ctx->stream>>>(F->n, d->base, d->cls, d->table, to, from, damp); break;
  }
  KCHECK(ctx);
  if (side) UG_TRY(side_end(ctx, xs));
  return 0;
}

int trc_restrict(uggpu_ctx *ctx, Level *F, Level *C, double *to, const double *from, Damp damp, bool fuse, const SellMat *Ac, double *tout, double *czero, Damp sdamp,
                 const HaloK &hk, int *done)
{
  *done = 0;
  if (!trc_enabled(C->n)) return 0;
  const bool comm = hk.flag != nullptr;
  const bool fpart = ctx->comm && F->partitioned;
  TrcData *d = nullptr;
  UG_TRY(trc_ensure(ctx, &F->R, C, true, fpart ? F->n : -1, (comm && C->partitioned) ? halo_snd_bits(C) : nullptr, comm, &d));
  if (!d) return 0;
  *done = 1;
  cudaStream_t xs = ctx->stream;
  const bool side = d->nx > 0 || comm;
  if (side) UG_TRY(side_begin(ctx, &xs));
  const int xblocks = (d->nx + 127) / 128 > 0 ? (d->nx + 127) / 128 : 1;
  const int nown = fpart ? F->n : 0x7fffffff;
  const double *diag = fuse ? Ac->diag : nullptr;
#define XR(BSV, FV) { if (comm) k_restrict_xrows<BSV, FV, true><<<xblocks, 128, 0, xs>>>(view(F->R), d->xrows, d->nx, nown, C->vnclass, C->skip, to, from, damp, diag, C->vclass, tout, czero, sdamp, ctx->derr, hk); \
                      else k_restrict_xrows<BSV, FV, false><<<xblocks, 128, 0, xs>>>(view(F->R), d->xrows, d->nx, nown, C->vnclass, C->skip, to, from, damp, diag, C->vclass, tout, czero, sdamp, ctx->derr, hk); }
#define XRB(BSV) { if (fuse) XR(BSV, true) else XR(BSV, false) }
  if (side) { switch (F->bs) { case 1: XRB(1) break; case 2: XRB(2) break; default: XRB(3) break; } KCHECK(ctx); }
#undef XRB
#undef XR
  const int blocks = (C->n + TRC_THREADS - 1) / TRC_THREADS;
#define RC(BSV) { if (fuse) k_restrict_cls<BSV, true><<<blocks, TRC_THREADS, 0, ctx->stream>>>(C->n, d->base, d->cls, d->table, to, from, damp, diag, tout, czero, sdamp, ctx->derr); \
                  else k_restrict_cls<BSV, false><<<blocks, TRC_THREADS, 0, ctx->stream>>>(C->n, d->base, d->cls, d->table, to, from, damp, diag, tout, czero, sdamp, ctx->derr); }
  if (blocks > 0) { switch (F->bs) { case 1: RC(1) break; case 2: RC(2) break; default: RC(3) break; } KCHECK(ctx); }
#undef RC
  if (side) UG_TRY(side_end(ctx, xs));
  return 0;
}
