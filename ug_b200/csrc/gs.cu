// gs.cu -- the Gauss-Seidel family and the ILU smoother of np/algebra/ugiter.cc on the device:
//   l_lgs :412  v = (D+L)^-1 d        l_ugs :735  v = (D+U)^-1 d        l_lsor :1343 / l_usor :1563  the same with relaxation
// and the smoother classes built on them (np/procs/iter.cc: gs :1039, sgs :1392, sor :4744/:4786).
//
// The reference solves the triangle row by row in VINDEX order; row r needs the NEW values of the rows on the solved side
// that it is connected to.  Here the rows are scheduled by dependency level (level 0: no such connection; level k: all of
// them on levels < k) -- every row still adds exactly the reference's terms in the reference's order (VSTART->MNEXT, the
// other side and inactive columns skipped), so the result is bit-identical, but all rows of a level run in parallel.
//
// PreProcess (uggpu_gs_preprocess, the analogue of GSPreProcess iter.cc:1003: l_setindex) builds, per direction:
//   * the levels, by a topological sweep over the (structurally symmetric: CONNECTION = MATRIX pair, gm/gm.h:653) pattern;
//   * a SECOND copy of the triangle, SELL-32 again, with the rows permuted into level order (levels padded to whole
//     slices) and each row holding [diagonal block, entries of the solved side in list order]: a warp reads a level's rows
//     with the same coalesced 128/256-byte loads as the SpMV kernels; column indices keep the original numbering.
// The solve is ONE persistent, cooperatively launched kernel per sweep with a static slice assignment (warp w owns slices w, w + W,
// ...); a slice waits for the slices its rows read from (point-to-point completion words) or, on schedules with more than 32
// dependencies per slice, for the whole previous level (per-SM counters + mailboxes) -- see "the solve" below.  A wait that exceeds
// 20 s sets the device error word instead of hanging.
//
// ILU (second half of this file): uggpu_dmatcopy / uggpu_l_ilubthdecomp / uggpu_l_luiter = class `ilu` (np/procs/iter.cc:5385-5525):
// the decomposition runs left-looking on the dependency levels of the lower triangle, one launch per level; l_luiter is two more
// modes of the solve kernel.
#include <cub/device/device_radix_sort.cuh>       // before uggpu_internal.h: its SLICE macro is an identifier inside cub
#include <cub/device/device_scan.cuh>

#include "uggpu_internal.h"

#include <cstdlib>
#include <vector>

#define TRI_THREADS 256

struct TriSched {
  int n = 0;                      // rows of the level
  int nT = 0;                     // rows of the schedule (levels padded to multiples of 32)
  int nlev = 0;
  SellMat T;                      // permuted triangle
  int32_t *perm = nullptr;        // [nT] schedule position -> row, -1 = padding
  int32_t *slice_level = nullptr; // [nT/32]
  int32_t *level_slices = nullptr;// [nlev]
  int32_t *level_first = nullptr; // [nlev] first slice of every level
  unsigned int *counters = nullptr;   // [nlev * 256 * 8]: finished slices per level and SM slot, one 32-byte sector each (k_trisolve)
  unsigned long long *mail = nullptr; // [256 * 16]: one progress word per SM slot, 128 bytes apart
  unsigned int epoch = 0;             // launches on this schedule so far
  // point-to-point mode: per slice the (at most TRI_DEPS) slices it reads from, and one completion word per slice
  int32_t *deps = nullptr;            // [nT/32 * 32], -1 = none
  unsigned int *sdone = nullptr;      // [nT/32]: epoch of the launch that finished the slice
  bool p2p_ok = false;
  std::vector<int> h_poff, h_lsl;     // host copies: first schedule row and slices of every level (the ILU decomposition launches per level)
};

static int tri_free(uggpu_ctx *ctx, TriSched *&S)
{
  if (!S) return 0;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  sell_free(ctx, &S->T);
  if (S->perm) dfree(ctx, S->perm, (size_t)S->nT);
  if (S->slice_level) dfree(ctx, S->slice_level, (size_t)S->nT / 32);
  if (S->level_slices) dfree(ctx, S->level_slices, (size_t)S->nlev);
  if (S->level_first) dfree(ctx, S->level_first, (size_t)S->nlev);
  if (S->counters) dfree(ctx, S->counters, (size_t)S->nlev * 256 * 8);
  if (S->mail) dfree(ctx, S->mail, (size_t)256 * 16);
  if (S->deps) dfree(ctx, S->deps, (size_t)S->nT);
  if (S->sdone) dfree(ctx, S->sdone, (size_t)S->nT / 32);
  delete S;
  S = nullptr;
  return 0;
}

int sell_free_schedules(uggpu_ctx *ctx, SellMat *m)
{
  UG_TRY(tri_free(ctx, m->tri[0]));
  UG_TRY(tri_free(ctx, m->tri[1]));
  return 0;
}

// ---- schedule construction ----------------------------------------------------------------------------------------------
// solved side of row r in direction dir: dir 0 (lower) columns c < r, dir 1 (upper) columns c > r
__device__ __forceinline__ bool solved_side(int dir, int r, int c) { return dir == 0 ? c < r : c > r; }

// indeg[r] = active connections of r on the solved side; rows without any (and all inactive rows) form level 0
__global__ void k_tri_indeg(SellView A, const uint8_t *__restrict__ vclass, int dir, int *__restrict__ indeg, int *__restrict__ lev,
                            int *__restrict__ frontier, int *__restrict__ count)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  int deg = 0;
  if (vclass[r] >= 3) {
    const int len = A.rowlen[r];
    const ColIter ci = col_iter(A, r);
    for (int j = 1; j < len; j++) {
      const int c = col_at(ci, j);
      if (solved_side(dir, r, c) && vclass[c] >= 3) deg++;
    }
  }
  indeg[r] = deg;
  lev[r] = deg == 0 ? 0 : -1;
  if (deg == 0) frontier[atomicAdd(count, 1)] = r;
}

// rows finished on level `cur` release the rows that list them on their solved side (= the rows they list on the OTHER side)
__global__ void k_tri_relax(SellView A, const uint8_t *__restrict__ vclass, int dir, const int *__restrict__ fin, int nin, int cur,
                            int *__restrict__ indeg, int *__restrict__ lev, int *__restrict__ fout, int *__restrict__ count, int *err)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nin) return;
  const int r = fin[i];
  if (vclass[r] < 3) return;                 // inactive rows are nobody's dependency
  const int len = A.rowlen[r];
  const ColIter ci = col_iter(A, r);
  for (int j = 1; j < len; j++) {
    const int c = col_at(ci, j);
    if (c == r || solved_side(dir, r, c) || vclass[c] < 3) continue;
    const int old = atomicSub(&indeg[c], 1);
    if (old == 1) { lev[c] = cur + 1; fout[atomicAdd(count, 1)] = c; }
    else if (old <= 0) atomicExch(err, UGGPU_ERROR);      // c does not list r: pattern not structurally symmetric
  }
}

__global__ void k_tri_iota(int n, int *__restrict__ a)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

__global__ void k_tri_place(int n, const int *__restrict__ key, const int *__restrict__ row, const int *__restrict__ start, const int *__restrict__ poff,
                            int32_t *__restrict__ perm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = key[i];
  perm[poff[l] + (i - start[l])] = row[i];
}

// entries of schedule row p: the diagonal block + the active entries of the solved side (0 for padding and inactive rows)
__global__ void k_tri_len(int nT, const int32_t *__restrict__ perm, SellView A, const uint8_t *__restrict__ vclass, int dir, int64_t *__restrict__ len)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nT) return;
  const int r = perm[p];
  int64_t n = 0;
  if (r >= 0 && vclass[r] >= 3) {
    n = 1;
    const int rl = A.rowlen[r];
    const ColIter ci = col_iter(A, r);
    for (int j = 1; j < rl; j++) {
      const int c = col_at(ci, j);
      if (solved_side(dir, r, c) && vclass[c] >= 3) n++;
    }
  }
  len[p] = n;
}

__global__ void k_tri_fill(int nT, int bb, const int32_t *__restrict__ perm, SellView A, const uint8_t *__restrict__ vclass, int dir,
                           const int64_t *__restrict__ rowptr, int32_t *__restrict__ ccol, double *__restrict__ cval)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nT) return;
  const int r = perm[p];
  if (r < 0 || vclass[r] < 3) return;
  const int rl = A.rowlen[r], lane = r & 31;
  const int64_t sp = slice_off(A, r >> 5);
  const ColIter ci = col_iter(A, r);
  int64_t o = rowptr[p];
  for (int j = 0; j < rl; j++) {
    const int c = col_at(ci, j);
    if (j > 0 && !(solved_side(dir, r, c) && vclass[c] >= 3)) continue;
    ccol[o] = c;
    for (int k = 0; k < bb; k++) cval[o * bb + k] = A.val[(sp + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane];
    o++;
  }
}

// pos[row] = schedule position
__global__ void k_tri_pos(int nT, const int32_t *__restrict__ perm, int32_t *__restrict__ pos)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nT && perm[p] >= 0) pos[perm[p]] = p;
}

// deps[s*32 .. +32): the distinct schedule slices the rows of slice s read from (one warp per slice, a 64-slot hash set in shared
// memory); more than 32 of them -> *overflow (the schedule then keeps to whole-level waits)
__global__ void __launch_bounds__(256) k_tri_deps(SellView T, const int32_t *__restrict__ pos, int32_t *__restrict__ deps, int *overflow)
{
  __shared__ int set[8][64];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nsl = T.n >> 5;
  if (s >= nsl) return;
  set[wib][lane] = -1; set[wib][lane + 32] = -1;
  __syncwarp();
  const int p = s * 32 + lane;
  const int len = T.rowlen[p];
  const ColIter ci = col_iter(T, p);
  bool over = false;
  for (int j = 1; j < len; j++) {
    const int ds = pos[col_at(ci, j)] >> 5;
    int h = ds & 63, probe = 0;
    for (; probe < 64; probe++) {
      const int old = atomicCAS(&set[wib][h], -1, ds);
      if (old == -1 || old == ds) break;
      h = (h + 1) & 63;
    }
    if (probe == 64) over = true;
  }
  __syncwarp();
  int base = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int val = set[wib][half * 32 + lane];
    const unsigned m = __ballot_sync(0xffffffffu, val >= 0);
    const int idx = base + __popc(m & ((1u << lane) - 1u));
    if (val >= 0) { if (idx < 32) deps[(size_t)s * 32 + idx] = val; else over = true; }
    base += __popc(m);
  }
  if (__any_sync(0xffffffffu, over) && lane == 0) atomicExch(overflow, 1);
}

static int tri_build(uggpu_ctx *ctx, Level *L, const SellMat *A, int dir, TriSched **out)
{
  cudaStream_t st = ctx->stream;
  const int n = L->n;
  TriSched *S = new TriSched();
  S->n = n;
  int *indeg = nullptr, *lev = nullptr, *fa = nullptr, *fb = nullptr, *d_count = nullptr;
  int rc = 0;
  std::vector<int> level_rows;          // rows per level
#define TB(expr) do { if ((rc = (expr)) != 0) goto fail; } while (0)
#define TC(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { rc = uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); goto fail; } } while (0)
  {
    TB(dalloc(ctx, &indeg, (size_t)n)); TB(dalloc(ctx, &lev, (size_t)n));
    TB(dalloc(ctx, &fa, (size_t)n)); TB(dalloc(ctx, &fb, (size_t)n)); TB(dalloc(ctx, &d_count, 1));
    const SellView Av = view(*A);
    const int blocks = (n + 255) / 256;
    int h_count = 0;
    TC(cudaMemsetAsync(d_count, 0, sizeof(int), st));
    k_tri_indeg<<<blocks, 256, 0, st>>>(Av, L->vclass, dir, indeg, lev, fa, d_count);
    ctx->launches++;
    TC(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    TC(cudaStreamSynchronize(st));
    int64_t total = 0;
    int cur = 0;
    while (h_count > 0) {
      level_rows.push_back(h_count);
      total += h_count;
      const int nin = h_count;
      TC(cudaMemsetAsync(d_count, 0, sizeof(int), st));
      k_tri_relax<<<(nin + 255) / 256, 256, 0, st>>>(Av, L->vclass, dir, fa, nin, cur, indeg, lev, fb, d_count, ctx->derr);
      ctx->launches++;
      TC(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
      TC(cudaStreamSynchronize(st));
      int *sw = fa; fa = fb; fb = sw;
      cur++;
    }
    TB(check_device_error(ctx));
    if (total != n) { rc = uggpu_fail(UGGPU_ERROR, "Gauss-Seidel schedule: %lld of %d rows could be ordered (the pattern is not structurally symmetric)", (long long)total, n); goto fail; }
    S->nlev = (int)level_rows.size();
    // rows sorted by (level, row): stable radix sort of the row numbers by level
    int *key_out = fa, *row_in = fb, *row_out = indeg;      // reuse: fa/fb/indeg are free now
    k_tri_iota<<<blocks, 256, 0, st>>>(n, row_in);
    ctx->launches++;
    int bits = 1;
    while ((1 << bits) < S->nlev && bits < 31) bits++;
    size_t tmp_bytes = 0;
    TC(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, lev, key_out, row_in, row_out, n, 0, bits, st));
    void *tmp = nullptr;
    TB(dev_alloc(ctx, &tmp, tmp_bytes ? tmp_bytes : 1));
    {
      cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, lev, key_out, row_in, row_out, n, 0, bits, st);
      ctx->launches++;
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      dev_free(ctx, tmp, tmp_bytes ? tmp_bytes : 1);
      if (e != cudaSuccess) { rc = uggpu_fail(UGGPU_CUDA_ERROR, "radix sort: %s", cudaGetErrorString(e)); goto fail; }
    }
    // level starts in the sorted order (= prefix sums of level_rows) and in the padded schedule
    std::vector<int> start(S->nlev), poff(S->nlev), lsl(S->nlev);
    int64_t acc = 0, pacc = 0;
    for (int l = 0; l < S->nlev; l++) {
      start[l] = (int)acc; poff[l] = (int)pacc;
      lsl[l] = (level_rows[l] + 31) / 32;
      acc += level_rows[l]; pacc += (int64_t)lsl[l] * 32;
    }
    if (pacc > 2147483647LL - 64) { rc = uggpu_fail(UGGPU_ERROR, "Gauss-Seidel schedule too long"); goto fail; }
    S->nT = (int)pacc;
    S->h_poff = poff; S->h_lsl = lsl;
    const int nslT = S->nT / 32;
    std::vector<int32_t> sl((size_t)nslT);
    for (int l = 0, s = 0; l < S->nlev; l++) for (int k = 0; k < lsl[l]; k++) sl[s++] = l;
    int *d_start = nullptr, *d_poff = nullptr;
    TB(dalloc(ctx, &d_start, (size_t)S->nlev)); TB(dalloc(ctx, &d_poff, (size_t)S->nlev));
    TB(dalloc(ctx, &S->perm, (size_t)S->nT)); TB(dalloc(ctx, &S->slice_level, (size_t)nslT));
    TB(dalloc(ctx, &S->level_slices, (size_t)S->nlev)); TB(dalloc(ctx, &S->counters, (size_t)S->nlev * 256 * 8)); TB(dalloc(ctx, &S->mail, (size_t)256 * 16));
    TB(dalloc(ctx, &S->level_first, (size_t)S->nlev));
    TC(cudaMemsetAsync(S->counters, 0, sizeof(unsigned int) * (size_t)S->nlev * 256 * 8, st));
    TC(cudaMemsetAsync(S->mail, 0, sizeof(unsigned long long) * (size_t)256 * 16, st));
    TC(cudaMemcpyAsync(d_start, start.data(), sizeof(int) * S->nlev, cudaMemcpyHostToDevice, st));
    TC(cudaMemcpyAsync(d_poff, poff.data(), sizeof(int) * S->nlev, cudaMemcpyHostToDevice, st));
    TC(cudaMemcpyAsync(S->level_slices, lsl.data(), sizeof(int) * S->nlev, cudaMemcpyHostToDevice, st));
    std::vector<int32_t> lfirst(S->nlev);
    for (int l = 0; l < S->nlev; l++) lfirst[l] = poff[l] / 32;          // first slice of the level
    TC(cudaMemcpyAsync(S->level_first, lfirst.data(), sizeof(int32_t) * S->nlev, cudaMemcpyHostToDevice, st));
    TC(cudaMemcpyAsync(S->slice_level, sl.data(), sizeof(int32_t) * nslT, cudaMemcpyHostToDevice, st));
    TC(cudaMemsetAsync(S->perm, 0xff, sizeof(int32_t) * (size_t)S->nT, st));
    k_tri_place<<<blocks, 256, 0, st>>>(n, key_out, row_out, d_start, d_poff, S->perm);
    ctx->launches++;
    TC(cudaStreamSynchronize(st));
    dfree(ctx, d_start, (size_t)S->nlev); dfree(ctx, d_poff, (size_t)S->nlev);
    // the permuted triangle as CSR, then SELL
    int64_t *d_len = nullptr, *d_rp = nullptr;
    int32_t *ccol = nullptr; double *cval = nullptr;
    TB(dalloc(ctx, &d_len, (size_t)S->nT + 1)); TB(dalloc(ctx, &d_rp, (size_t)S->nT + 1));
    TC(cudaMemsetAsync(d_len, 0, sizeof(int64_t) * ((size_t)S->nT + 1), st));
    const int tblocks = (S->nT + 255) / 256;
    k_tri_len<<<tblocks, 256, 0, st>>>(S->nT, S->perm, Av, L->vclass, dir, d_len);
    ctx->launches++;
    tmp_bytes = 0;
    TC(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_len, d_rp, S->nT + 1, st));
    TB(dev_alloc(ctx, &tmp, tmp_bytes ? tmp_bytes : 1));
    int64_t nnzT = 0;
    {
      cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_len, d_rp, S->nT + 1, st);
      ctx->launches++;
      if (e == cudaSuccess) e = cudaMemcpyAsync(&nnzT, d_rp + S->nT, sizeof(int64_t), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      dev_free(ctx, tmp, tmp_bytes ? tmp_bytes : 1);
      if (e != cudaSuccess) { rc = uggpu_fail(UGGPU_CUDA_ERROR, "scan: %s", cudaGetErrorString(e)); goto fail; }
    }
    dfree(ctx, d_len, (size_t)S->nT + 1);
    TB(dalloc(ctx, &ccol, (size_t)(nnzT > 0 ? nnzT : 1))); TB(dalloc(ctx, &cval, (size_t)(nnzT > 0 ? nnzT : 1) * A->bb));
    k_tri_fill<<<tblocks, 256, 0, st>>>(S->nT, A->bb, S->perm, Av, L->vclass, dir, d_rp, ccol, cval);
    ctx->launches++;
    rc = sell_from_device_csr(ctx, S->nT, A->bb, d_rp, ccol, cval, &S->T);
    cudaStreamSynchronize(st);
    dfree(ctx, d_rp, (size_t)S->nT + 1); dfree(ctx, ccol, (size_t)(nnzT > 0 ? nnzT : 1)); dfree(ctx, cval, (size_t)(nnzT > 0 ? nnzT : 1) * A->bb);
    if (rc) goto fail;
    // point-to-point dependencies of every slice (pos = inverse of perm; lev[] is free by now and has n entries)
    if (!getenv("UGGPU_GS_LEVELS")) {
      int32_t *pos = lev;
      int *d_over = d_count;
      TB(dalloc(ctx, &S->deps, (size_t)S->nT)); TB(dalloc(ctx, &S->sdone, (size_t)S->nT / 32));
      TC(cudaMemsetAsync(S->deps, 0xff, sizeof(int32_t) * (size_t)S->nT, st));
      TC(cudaMemsetAsync(S->sdone, 0, sizeof(unsigned int) * (size_t)S->nT / 32, st));
      TC(cudaMemsetAsync(d_over, 0, sizeof(int), st));
      k_tri_pos<<<(S->nT + 255) / 256, 256, 0, st>>>(S->nT, S->perm, pos);
      ctx->launches++;
      k_tri_deps<<<(S->nT / 32 + 7) / 8, 256, 0, st>>>(view(S->T), pos, S->deps, d_over);
      ctx->launches++;
      int over = 1;
      TC(cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, st));
      TC(cudaStreamSynchronize(st));
      S->p2p_ok = over == 0;
    }
  }
  dfree(ctx, indeg, (size_t)n); dfree(ctx, lev, (size_t)n); dfree(ctx, fa, (size_t)n); dfree(ctx, fb, (size_t)n); dfree(ctx, d_count, 1);
  *out = S;
  return 0;
fail:
  if (indeg) dfree(ctx, indeg, (size_t)n);
  if (lev) dfree(ctx, lev, (size_t)n);
  if (fa) dfree(ctx, fa, (size_t)n);
  if (fb) dfree(ctx, fb, (size_t)n);
  if (d_count) dfree(ctx, d_count, 1);
  tri_free(ctx, S);
  return rc;
#undef TB
#undef TC
}

extern "C" int uggpu_gs_preprocess(uggpu_ctx *ctx, int level, int M)
{
  Level *L = get_level(ctx, level);
  SellMat *A = get_mat(ctx, level, M);
  if (!L || !A) return UGGPU_DESC_MISMATCH;
  if (ctx->comm && L->partitioned) return uggpu_fail(UGGPU_ERROR, "Gauss-Seidel smoothers run on one GPU (level %d is partitioned)", level);
  if (L->n == 0) return 0;
  for (int dir = 0; dir < 2; dir++)
    if (!A->tri[dir]) UG_TRY(tri_build(ctx, L, A, dir, &A->tri[dir]));
  return 0;
}

extern "C" int uggpu_gs_levels(uggpu_ctx *ctx, int level, int M, int *lower, int *upper)
{
  SellMat *A = get_mat(ctx, level, M);
  if (!A) return UGGPU_DESC_MISMATCH;
  if (lower) *lower = A->tri[0] ? A->tri[0]->nlev : 0;
  if (upper) *upper = A->tri[1] ? A->tri[1]->nlev : 0;
  return 0;
}

// ---- the solve -------------------------------------------------------------------------------------------------------------
// One persistent kernel per sweep, launched cooperatively (all warps co-resident) with a STATIC assignment: warp w owns slices
// w, w + W, w + 2W, ... of the schedule.  Synchronisation words of a schedule (never reset: every launch has an epoch number):
//   done[(lv * TRI_SLOTS + slot) * 8]   finished slices of level lv, one counter per SM slot (smid mod TRI_SLOTS), one 32-byte sector
//                                       each: a warp that finishes a slice adds 1 with a fire-and-forget reduction -- no burst of
//                                       same-address atomics at the end of a level;
//   mail[slot * 16] (64 bit)            one mailbox per SM: epoch * (nlev + 1) + number of completed levels.
// The warp that owns the FIRST slice of level lv + 1 is the monitor of level lv: instead of waiting it sums the level's counters
// (8 acquire loads per lane) until they reach the level's slice count, then raises every mailbox (atomicMax).  All other warps
// poll only the mailbox of their own SM, rarely while they are several levels away.
// History at 257^3 (769 levels per sweep): every warp polling the level counter 54 us per level; mailboxes + tickets 11 us.
#define TRI_SLOTS 256
#define TRI_PRE 8            // column indices of a row held in registers while the warp waits

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int smid() { unsigned int v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); return v; }

struct TriArgs {
  const int32_t *perm, *slice_level, *level_slices, *level_first;
  unsigned int *done;
  unsigned long long *mail;
  int nlev, nsl;
  unsigned int epoch;         // 1, 2, 3, ... per launch on this schedule
  const int32_t *deps;        // point-to-point mode (P2P)
  unsigned int *sdone;
};

// SOR (the solve's mode): 0 = l_lgs / l_ugs, 1 = l_lsor / l_usor (scalar rows: omega*(d-sum)/diag ugiter.cc:1400; block rows:
// solve, then v_i *= omega_i :1556), 2 = lower sweep of l_luiter (Diag(L) = I: v = d - sum, ugiter.cc:4490,4648), 3 = upper sweep
// of l_luiter (right-hand side = v itself, stored inverse diagonal applied by multiplication :4510, SolveInverseSmallBlock
// block.cc:225)
// P2P: a slice waits for the slices it reads from (one completion word per slice) instead of the whole previous level -- no
// monitor, no mailbox hop, no tail at the end of every level; available when every slice reads from at most 32 others.
template <int BS, int SOR, bool P2P>
__global__ void __launch_bounds__(TRI_THREADS) k_trisolve(SellView T, TriArgs a, double *v, const double *__restrict__ d, Damp omega, int *err)
{
  constexpr int BB = BS * BS;
  const int lane = threadIdx.x & 31;
  const int W = (int)((gridDim.x * blockDim.x) >> 5);
  const unsigned int slot = smid() & (TRI_SLOTS - 1);
  const unsigned long long *const mybox = a.mail + (size_t)slot * 16;
  const unsigned long long base = (unsigned long long)a.epoch * (unsigned long long)(a.nlev + 1);
  int known = 0;               // levels 0 .. known-1 are complete, as far as this warp has seen
  for (int s = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); s < a.nsl; s += W) {
    const int lv = a.slice_level[s];
    const int p = s * 32 + lane;
    if (s + W < a.nsl && lane < 3)          // the warp's NEXT slice: its direct-indexed lines into L2
      prefetch_l2(lane == 0 ? (const void *)(a.perm + (size_t)(s + W) * 32) : lane == 1 ? (const void *)(T.rowlen + (size_t)(s + W) * 32) : (const void *)(a.slice_level + s + W));
    const int r = a.perm[p];
    const int len = r >= 0 ? (int)T.rowlen[p] : 0;
    const int64_t sp = slice_off(T, s);
    const int w = slice_width(T, s, sp);
    // everything that does not depend on other rows is fetched BEFORE the wait: the slice's values into L2, the row's column
    // indices, diagonal block and right-hand side into registers
    if (lv > known) {
      const char *vb = reinterpret_cast<const char *>(T.val + sp * BB);
      const int vlines = w * BB * 2;                        // 256 bytes per component and slice column
      for (int l = lane; l < vlines; l += 32) prefetch_l2(vb + (size_t)l * 128);
    }
    const ColIter ci = col_iter(T, p);
    const double *__restrict__ vp = T.val + sp * BB + lane;
    int cj[TRI_PRE];
#pragma unroll
    for (int j = 1; j <= TRI_PRE; j++) cj[j - 1] = j < len ? col_at(ci, j) : 0;
    double dg[BB], rhs[BS];
#pragma unroll
    for (int k = 0; k < BB; k++) dg[k] = len > 0 ? __ldg(vp + (size_t)k * 32) : 1.0;
#pragma unroll
    for (int i = 0; i < BS; i++) rhs[i] = len > 0 ? (SOR == 3 ? __ldcg(v + (size_t)r * BS + i) : d[(size_t)r * BS + i]) : 0.0;
    if (P2P) {
      // every lane watches one of the slices this one reads from; all of them poll in parallel, so a finished dependency costs
      // one L2 round trip, not one per stage.  The acquire loads + the warp vote order the gathers below behind the writers'
      // release (fence + flag store) without another fence.
      const int dep = a.deps[(size_t)s * 32 + lane];
      bool ok = dep < 0 || ld_acquire_u32(a.sdone + dep) == a.epoch;
      if (!__all_sync(0xffffffffu, ok)) {
        unsigned long long t0 = 0, t1;
        int spins = 0;
        unsigned int ns = 32;
        do {
          if (!ok) {
            __nanosleep(ns);
            if (ns < 256) ns <<= 1;
            ok = ld_acquire_u32(a.sdone + dep) == a.epoch;
            if (++spins == 1024) {
              spins = 0;
              if (*reinterpret_cast<volatile int *>(err)) ok = true;          // an earlier wait already failed: do not wait again
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
              if (t0 == 0) t0 = t1;
              else if (t1 - t0 > 20000000000ull) { atomicExch(err, UGGPU_CUDA_ERROR); ok = true; }
            }
          }
        } while (!__all_sync(0xffffffffu, ok));
      }
    } else if (lv > known) {
      unsigned long long t0 = 0, t1;
      int spins = 0;
      if (s == a.level_first[lv]) {
        // monitor of level lv - 1
        const unsigned int need = (unsigned int)a.level_slices[lv - 1];
        const unsigned int want = a.epoch * need;                            // modulo 2^32, like the counters: every launch adds `need`
        const unsigned int *cnt = a.done + (size_t)(lv - 1) * TRI_SLOTS * 8;
        for (;;) {
          unsigned int sum = 0;
#pragma unroll
          for (int q = 0; q < TRI_SLOTS / 32; q++) sum += ld_acquire_u32(cnt + (size_t)(lane + 32 * q) * 8);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (sum == want) break;
          int give_up = 0;
          if (lane == 0 && ++spins == 256) {
            spins = 0;
            if (*reinterpret_cast<volatile int *>(err)) give_up = 1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            else if (t1 - t0 > 20000000000ull) { atomicExch(err, UGGPU_CUDA_ERROR); give_up = 1; }
          }
          if (__shfl_sync(0xffffffffu, give_up, 0)) break;
        }
        __threadfence();
        __syncwarp();
        for (int q = lane; q < TRI_SLOTS; q += 32) atomicMax(a.mail + (size_t)q * 16, base + (unsigned long long)lv);
        known = lv;
      } else {
        int cur = known;
        if (lane == 0) {
          for (;;) {
            const unsigned long long m = ld_acquire_u64(mybox);
            cur = m > base ? (int)(m - base) : 0;
            if (cur >= lv) break;
            // far from its turn a warp polls rarely (a level takes a few us); next in line it polls tightly
            const unsigned int dist = (unsigned int)(lv - cur);
            __nanosleep(dist > 1 ? min((dist - 1) * 1500u, 12000u) : 40u);
            if (++spins == 64) {
              spins = 0;
              if (*reinterpret_cast<volatile int *>(err)) break;          // an earlier wait already failed: do not wait again
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
              if (t0 == 0) t0 = t1;
              else if (t1 - t0 > 20000000000ull) { atomicExch(err, UGGPU_CUDA_ERROR); break; }
            }
          }
        }
        known = max(known, __shfl_sync(0xffffffffu, cur, 0));
      }
    }
    if (r >= 0) {
      double sol[BS];
      if (len == 0) {                                        // VCLASS < ACTIVE_CLASS: v = 0 (ugiter.cc:447)
#pragma unroll
        for (int i = 0; i < BS; i++) sol[i] = 0.0;
      } else {
        double acc[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) acc[i] = 0.0;
        // entry j: the values were prefetched into L2, the operand was written by another SM during this launch (L2, not L1)
        auto term = [&](int j, int c) {
          double m[BB], wv[BS];
#pragma unroll
          for (int k = 0; k < BB; k++) m[k] = __ldg(vp + ((size_t)j * BB + k) * 32);
#pragma unroll
          for (int i = 0; i < BS; i++) wv[i] = __ldcg(v + (size_t)c * BS + i);
#pragma unroll
          for (int i = 0; i < BS; i++) {
            double t = m[i * BS] * wv[0];
#pragma unroll
            for (int q = 1; q < BS; q++) t = t + m[i * BS + q] * wv[q];
            acc[i] += t;
          }
        };
#pragma unroll
        for (int j = 1; j <= TRI_PRE; j++)
          if (j < len) term(j, cj[j - 1]);
        for (int j = TRI_PRE + 1; j < len; j++) term(j, col_at(ci, j));
        if (BS == 1) {
          if (SOR == 1) sol[0] = omega.a[0] * (rhs[0] - acc[0]) / dg[0];
          else if (SOR == 2) sol[0] = rhs[0] - acc[0];
          else if (SOR == 3) sol[0] = (rhs[0] - acc[0]) * dg[0];
          else sol[0] = (rhs[0] - acc[0]) / dg[0];
        } else {
#pragma unroll
          for (int i = 0; i < BS; i++) rhs[i] = rhs[i] - acc[i];
          // SolveSmallBlock (block.cc:104-142), same closed forms as solve_small_block in spmv.cu
          if (SOR == 2) {
#pragma unroll
            for (int i = 0; i < BS; i++) sol[i] = rhs[i];
          } else if (SOR == 3) {
#pragma unroll
            for (int i = 0; i < BS; i++) {
              double sum = 0.0;
#pragma unroll
              for (int q = 0; q < BS; q++) sum += dg[i * BS + q] * rhs[q];
              sol[i] = sum;
            }
          } else if (BS == 2) {
            double det = dg[0] * dg[3 % BB] - dg[1 % BB] * dg[2 % BB];
            if (det == 0.0) { atomicExch(err, UGGPU_SMALL_DIAG); det = 1.0; }
            det = 1.0 / det;
            sol[0] = (rhs[0] * dg[3 % BB] - rhs[1 % BS] * dg[1 % BB]) * det;
            sol[1 % BS] = (rhs[1 % BS] * dg[0] - rhs[0] * dg[2 % BB]) * det;
          } else {
            double M3div0 = dg[3 % BB] / dg[0];
            double M6div0 = dg[6 % BB] / dg[0];
            double aux = (dg[7 % BB] - M6div0 * dg[1 % BB]) / (dg[4 % BB] - M3div0 * dg[1 % BB]);
            sol[2 % BS] = (rhs[2 % BS] - M6div0 * rhs[0] - aux * (rhs[1 % BS] - M3div0 * rhs[0]))
                          / (dg[8 % BB] - M6div0 * dg[2 % BB] - aux * (dg[5 % BB] - M3div0 * dg[2 % BB]));
            sol[1 % BS] = (rhs[1 % BS] - dg[3 % BB] / dg[0] * rhs[0] - (dg[5 % BB] - M3div0 * dg[2 % BB]) * sol[2 % BS])
                          / (dg[4 % BB] - M3div0 * dg[1 % BB]);
            sol[0] = (rhs[0] - dg[1 % BB] * sol[1 % BS] - dg[2 % BB] * sol[2 % BS]) / dg[0];
          }
          if (SOR == 1) {
#pragma unroll
            for (int i = 0; i < BS; i++) sol[i] = sol[i] * omega.a[i];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < BS; i++) __stcg(v + (size_t)r * BS + i, sol[i]);
    }
    // publish: the warp's stores, then one release by lane 0 (the pattern of a grid barrier, per warp)
    __syncwarp();
    if (lane == 0) {
      __threadfence();
      if (P2P) *reinterpret_cast<volatile unsigned int *>(a.sdone + s) = a.epoch;
      else atomicAdd(a.done + ((size_t)lv * TRI_SLOTS + slot) * 8, 1u);
    }
  }
}

// co-resident blocks of a k_trisolve instantiation (cooperative launch)
template <int BS, int SOR, bool P2P>
static int tri_launch2(uggpu_ctx *ctx, const SellView &Tv, const TriArgs &a, double *v, const double *d, const Damp &om)
{
  static int per_sm = 0;       // per instantiation
  if (per_sm == 0) {
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trisolve<BS, SOR, P2P>, TRI_THREADS, 0));
    if (per_sm < 1) return uggpu_fail(UGGPU_CUDA_ERROR, "k_trisolve does not fit on an SM");
  }
  int blocks = (a.nsl + TRI_THREADS / 32 - 1) / (TRI_THREADS / 32);
  if (blocks > per_sm * ctx->sm_count) blocks = per_sm * ctx->sm_count;
  SellView tv = Tv; TriArgs aa = a; Damp o = om; int *err = ctx->derr;
  void *args[] = {&tv, &aa, &v, &d, &o, &err};
  CUDA_TRY(cudaLaunchCooperativeKernel((const void *)k_trisolve<BS, SOR, P2P>, dim3(blocks), dim3(TRI_THREADS), args, 0, ctx->stream));
  ctx->launches++;
  return 0;
}
template <int BS, int SOR>
static int tri_launch(uggpu_ctx *ctx, const SellView &Tv, const TriArgs &a, double *v, const double *d, const Damp &om)
{
  return a.deps ? tri_launch2<BS, SOR, true>(ctx, Tv, a, v, d, om) : tri_launch2<BS, SOR, false>(ctx, Tv, a, v, d, om);
}

// lu: 0 Gauss-Seidel / SOR (omega != NULL), 2 / 3 lower / upper sweep of l_luiter (mode of k_trisolve)
static int tri_solve(uggpu_ctx *ctx, int level, int M, int dir, double *v, const double *d, const double *omega, int lu = 0)
{
  Level *L = get_level(ctx, level);
  SellMat *A = get_mat(ctx, level, M);
  if (!L || !A) return UGGPU_DESC_MISMATCH;
  if (v == d && lu != 3) return uggpu_fail(UGGPU_DESC_MISMATCH, "Gauss-Seidel solve: result and right-hand side are the same vector");
  if (L->n == 0) return 0;
  if (!A->tri[dir]) UG_TRY(uggpu_gs_preprocess(ctx, level, M));
  TriSched *S = A->tri[dir];
  if (++S->epoch == 0xffffffffu) {       // the epoch arithmetic of the counters is modulo 2^32: start over well before it wraps
    CUDA_TRY(cudaMemsetAsync(S->counters, 0, sizeof(unsigned int) * (size_t)S->nlev * TRI_SLOTS * 8, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(S->mail, 0, sizeof(unsigned long long) * (size_t)TRI_SLOTS * 16, ctx->stream));
    if (S->sdone) CUDA_TRY(cudaMemsetAsync(S->sdone, 0, sizeof(unsigned int) * (size_t)S->nT / 32, ctx->stream));
    S->epoch = 1;
  }
  const TriArgs a{S->perm, S->slice_level, S->level_slices, S->level_first, S->counters, S->mail, S->nlev, S->nT / 32, S->epoch,
                  S->p2p_ok ? S->deps : nullptr, S->sdone};
  const Damp om = mkdamp(omega, L->bs);
  const SellView Tv = view(S->T);
  // algorithmic bytes: the triangle's entries, row lengths and permutation, d read, v written, gathered v once
  ProfScope ps(ctx, UGGPU_K_TRISOLVE, level, S->T.entry_bytes() + 6.0 * S->nT + 8.0 * L->bs * 3.0 * L->n);
  const int mode = lu ? lu : (omega ? 1 : 0);
#define TRI_CASE(BS_) \
  switch (mode) { case 0: return tri_launch<BS_, 0>(ctx, Tv, a, v, d, om); case 1: return tri_launch<BS_, 1>(ctx, Tv, a, v, d, om); \
                  case 2: return tri_launch<BS_, 2>(ctx, Tv, a, v, d, om); default: return tri_launch<BS_, 3>(ctx, Tv, a, v, d, om); }
  switch (L->bs) {
    case 1: TRI_CASE(1)
    case 2: TRI_CASE(2)
    default: TRI_CASE(3)
  }
#undef TRI_CASE
}

static int tri_entry(uggpu_ctx *ctx, int level, int v, int M, int d, int dir, const double *omega)
{
  double *vp = get_vec(ctx, level, v);
  const double *dp = get_vec(ctx, level, d);
  if (!vp || !dp) return UGGPU_DESC_MISMATCH;
  UG_TRY(tri_solve(ctx, level, M, dir, vp, dp, omega));
  return check_device_error(ctx);
}

extern "C" int uggpu_l_lgs(uggpu_ctx *ctx, int level, int v, int M, int d) { return tri_entry(ctx, level, v, M, d, 0, nullptr); }
extern "C" int uggpu_l_ugs(uggpu_ctx *ctx, int level, int v, int M, int d) { return tri_entry(ctx, level, v, M, d, 1, nullptr); }
extern "C" int uggpu_l_lsor(uggpu_ctx *ctx, int level, int v, int M, int d, const double *omega)
{
  if (!omega) return uggpu_fail(UGGPU_ERROR, "l_lsor: null omega");
  return tri_entry(ctx, level, v, M, d, 0, omega);
}
extern "C" int uggpu_l_usor(uggpu_ctx *ctx, int level, int v, int M, int d, const double *omega)
{
  if (!omega) return uggpu_fail(UGGPU_ERROR, "l_usor: null omega");
  return tri_entry(ctx, level, v, M, d, 1, omega);
}

// ---- ILU: l_ilubthdecomp ugiter.cc:2252 as class `ilu` calls it, l_luiter :4444 ----------------------------------
// The reference eliminates right-looking: row i (ascending) turns every entry (j,i), j > i, into the pivot M_ji * D_i^-1 and sends
// -pivot * M_ik into row j for the k > i of row i in row i's list order.  Seen from row j these are: for its lower neighbours i in
// ASCENDING order, the entries of row i in list order -- a left-looking sweep that needs rows i complete, i.e. exactly the
// dependency levels of the lower triangular solve.  One launch per level (the kernel boundary is the barrier: no spin-waits in a
// setup step), one thread per row; the row's own entries are updated in place in the matrix's SELL storage, with separate
// multiply and subtract like the reference's statements.  Connections are never created (no threshold, VCUSED = 0): what falls
// off the pattern goes into the beta-modification of the diagonal (scalar :2418, blocks :2598-2640).
template <int BS>
__global__ void __launch_bounds__(128) k_ilu_factor_level(SellView A, double *val, const uint8_t *__restrict__ vclass,
                                                          const int32_t *__restrict__ perm, int p0, int p1, Damp beta, int use_beta, int *err)
{
  constexpr int BB = BS * BS;
  const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= p1) return;
  const int r = perm[p];
  if (r < 0 || vclass[r] < 3) return;                       // L_VLOOP__CLASS(..., ACTIVE_CLASS): other rows stay as copied
  const int len = A.rowlen[r];
  const ColIter ci = col_iter(A, r);
  double *vr = val + slice_off(A, r >> 5) * BB + (r & 31);   // component k of entry j of this row: vr[(j*BB + k)*32]
  int last = -1;
  for (;;) {
    int i = 0x7fffffff, ji = -1;                            // the next lower neighbour in index order
    for (int j = 1; j < len; j++) {
      const int c = col_at(ci, j);
      if (c > last && c < r && c < i && vclass[c] >= 3) { i = c; ji = j; }
    }
    if (ji < 0) break;
    last = i;
    const int leni = A.rowlen[i];
    const ColIter cii = col_iter(A, i);
    const double *vi = val + slice_off(A, i >> 5) * BB + (i & 31);
    if (BS == 1) {
      const double pivot = vr[(size_t)ji * 32] * vi[0];     // vi[0]: the stored inverse diagonal of row i
      vr[(size_t)ji * 32] = pivot;
      if (pivot == 0.0) continue;
      for (int q = 1; q < leni; q++) {
        const int k = col_at(cii, q);
        if (!(k > i && vclass[k] >= 3)) continue;
        const double mik = vi[(size_t)q * 32];
        int t = -1;                                          // GetMatrix(vj, vk)
        for (int j = 0; j < len; j++) if (col_at(ci, j) == k) { t = j; break; }
        if (t >= 0) { const double pr = pivot * mik; vr[(size_t)t * 32] = vr[(size_t)t * 32] - pr; }
        else if (use_beta) { const double pr = beta.a[0] * fabs(pivot * mik); vr[0] = vr[0] + pr; }
      }
    } else {
      double inv[BB], pm[BB], pv[BB], cor[BB], rowsum[BS];
#pragma unroll
      for (int k = 0; k < BB; k++) { inv[k] = vi[(size_t)k * 32]; pv[k] = vr[((size_t)ji * BB + k) * 32]; }
      const bool pivzero = block_mul<BS>(pv, inv, pm);
#pragma unroll
      for (int k = 0; k < BB; k++) vr[((size_t)ji * BB + k) * 32] = pm[k];
      if (pivzero) continue;
#pragma unroll
      for (int l = 0; l < BS; l++) rowsum[l] = 0.0;
      for (int q = 1; q < leni; q++) {
        const int k = col_at(cii, q);
        if (!(k > i && vclass[k] >= 3)) continue;
        double elm[BB];
#pragma unroll
        for (int c = 0; c < BB; c++) elm[c] = vi[((size_t)q * BB + c) * 32];
        if (block_mul<BS>(pm, elm, cor)) continue;           // CorIsZero
        int t = -1;
        for (int j = 0; j < len; j++) if (col_at(ci, j) == k) { t = j; break; }
        if (t >= 0) {
#pragma unroll
          for (int c = 0; c < BB; c++) vr[((size_t)t * BB + c) * 32] = vr[((size_t)t * BB + c) * 32] - cor[c];
        } else {
#pragma unroll
          for (int l = 0; l < BS; l++)
#pragma unroll
            for (int m = 0; m < BS; m++) rowsum[l] += fabs(cor[l * BS + m]);      // normalisation factors are 1 without a rest vector
        }
      }
      if (!use_beta) continue;
      double dampf[BS];
#pragma unroll
      for (int m = 0; m < BS; m++) dampf[m] = 1.0 + beta.a[m] * rowsum[m];
#pragma unroll
      for (int m = 0; m < BS; m++)
#pragma unroll
        for (int l = 0; l < BS; l++) vr[(size_t)(m * BS + l) * 32] = vr[(size_t)(m * BS + l) * 32] * dampf[l];
    }
  }
  // the row's own diagonal: invert and store the inverse (:2367-2373 scalar, :2445-2452 blocks)
  if (BS == 1) {
    const double diag = vr[0];
    if (fabs(diag) < 2.220446049250313e-16 * 10.0 * 1e-20) { atomicExch(err, UGGPU_SMALL_DIAG); return; }   // SMALL_D*1e-20
    vr[0] = 1.0 / diag;
  } else {
    double dg[BB], inv[BB];
#pragma unroll
    for (int k = 0; k < BB; k++) dg[k] = vr[(size_t)k * 32];
    if (invert_small_block<BS>(dg, inv)) { atomicExch(err, UGGPU_SMALL_DIAG); return; }
#pragma unroll
    for (int k = 0; k < BB; k++) vr[(size_t)k * 32] = inv[k];
  }
}

// the values of the permuted triangle T after the matrix it was cut from changed (same entry selection and order as k_tri_fill)
__global__ void k_tri_refresh(int nT, int bb, const int32_t *__restrict__ perm, SellView A, const double *__restrict__ aval,
                              const uint8_t *__restrict__ vclass, int dir, SellView T, double *__restrict__ tval)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nT) return;
  const int r = perm[p];
  if (r < 0 || vclass[r] < 3) return;
  const int rl = A.rowlen[r], lane = r & 31;
  const int64_t sp = slice_off(A, r >> 5), spT = slice_off(T, p >> 5);
  const ColIter ci = col_iter(A, r);
  int o = 0;
  for (int j = 0; j < rl; j++) {
    const int c = col_at(ci, j);
    if (j > 0 && !(solved_side(dir, r, c) && vclass[c] >= 3)) continue;
    for (int k = 0; k < bb; k++) tval[(spT + (int64_t)o * 32) * bb + (int64_t)k * 32 + (p & 31)] = aval[(sp + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane];
    o++;
  }
}

// existing solve schedules of A carry their own copy of the values: bring them up to date after A's values changed
static int tri_refresh_all(uggpu_ctx *ctx, Level *L, SellMat *A)
{
  for (int dir = 0; dir < 2; dir++) {
    TriSched *T = A->tri[dir];
    if (!T) continue;
    k_tri_refresh<<<(T->nT + 255) / 256, 256, 0, ctx->stream>>>(T->nT, A->bb, T->perm, view(*A), A->val, L->vclass, dir, view(T->T), T->T.val);
    KCHECK(ctx);
    UG_TRY(sell_update_diag(ctx, &T->T));
  }
  return 0;
}

extern "C" int uggpu_dmatcopy(uggpu_ctx *ctx, int fl, int tl, int mode, int M, int A)
{
  if (mode != UGGPU_ALL_VECTORS) return uggpu_fail(UGGPU_ERROR, "dmatcopy: only ALL_VECTORS is supported");
  if (M == A) return uggpu_fail(UGGPU_DESC_MISMATCH, "dmatcopy: source and destination are the same matrix");
  for (int l = fl; l <= tl; l++) {
    Level *L = get_level(ctx, l);
    SellMat *src = get_mat(ctx, l, A);
    if (!L || !src) return UGGPU_DESC_MISMATCH;
    auto it = L->mats.find(M);
    if (it != L->mats.end()) {
      SellMat *dst = &it->second;
      if (dst->n == src->n && dst->bb == src->bb && dst->nnz == src->nnz && dst->padded == src->padded && dst->col_len == src->col_len && dst->fixed_w == src->fixed_w) {
        // same pattern (made by an earlier copy): the values only, also into the solve schedules M may have
        CUDA_TRY(cudaMemcpyAsync(dst->val, src->val, sizeof(double) * (size_t)src->padded * src->bb, cudaMemcpyDeviceToDevice, ctx->stream));
        UG_TRY(sell_drop_shared_values(ctx, dst));
        UG_TRY(sell_update_diag(ctx, dst));
        UG_TRY(tri_refresh_all(ctx, L, dst));
        continue;
      }
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      UG_TRY(sell_free(ctx, dst));
      L->mats.erase(it);
    }
    SellMat m;
    UG_TRY(sell_clone(ctx, src, &m));
    UG_TRY(sell_drop_shared_values(ctx, &m));      // the copy's code words no longer point into src's value tables
    L->mats[M] = m;
  }
  return 0;
}

extern "C" int uggpu_l_ilubthdecomp(uggpu_ctx *ctx, int level, int M, const double *beta)
{
  Level *L = get_level(ctx, level);
  SellMat *A = get_mat(ctx, level, M);
  if (!L || !A) return UGGPU_DESC_MISMATCH;
  if (ctx->comm && L->partitioned) return uggpu_fail(UGGPU_ERROR, "the ILU smoother runs on one GPU (level %d is partitioned)", level);
  if (L->n == 0) return 0;
  if (L->bs < 1 || L->bs > 3) return uggpu_fail(UGGPU_BLOCK_TOO_LARGE, "l_ilubthdecomp: block size %d", L->bs);
  UG_TRY(sell_drop_shared_values(ctx, A));         // the values are about to change in place
  // the dependency levels of the lower triangle (l_setindex + the pattern; values do not matter)
  if (!A->tri[0]) UG_TRY(tri_build(ctx, L, A, 0, &A->tri[0]));
  TriSched *S = A->tri[0];
  const SellView Av = view(*A);
  const Damp b = mkdamp(beta, L->bs);
  Damp bb0 = b;
  if (!beta) for (int i = 0; i < UGGPU_MAX_BS; i++) bb0.a[i] = 0.0;
  for (int lv = 0; lv < S->nlev; lv++) {
    const int p0 = S->h_poff[lv], p1 = p0 + 32 * S->h_lsl[lv];
    const int blocks = (p1 - p0 + 127) / 128;
    switch (L->bs) {
      case 1: k_ilu_factor_level<1><<<blocks, 128, 0, ctx->stream>>>(Av, A->val, L->vclass, S->perm, p0, p1, bb0, beta != nullptr, ctx->derr); break;
      case 2: k_ilu_factor_level<2><<<blocks, 128, 0, ctx->stream>>>(Av, A->val, L->vclass, S->perm, p0, p1, bb0, beta != nullptr, ctx->derr); break;
      default: k_ilu_factor_level<3><<<blocks, 128, 0, ctx->stream>>>(Av, A->val, L->vclass, S->perm, p0, p1, bb0, beta != nullptr, ctx->derr); break;
    }
    ctx->launches++;
  }
  CUDA_TRY(cudaGetLastError());
  UG_TRY(check_device_error(ctx));
  UG_TRY(sell_update_diag(ctx, A));
  // the solve schedules carry their own copy of the values: refresh what exists, then build what is missing (from the new values)
  UG_TRY(tri_refresh_all(ctx, L, A));
  for (int dir = 0; dir < 2; dir++)
    if (!A->tri[dir]) UG_TRY(tri_build(ctx, L, A, dir, &A->tri[dir]));
  return 0;
}

extern "C" int uggpu_l_luiter(uggpu_ctx *ctx, int level, int v, int M, int d)
{
  double *vp = get_vec(ctx, level, v);
  const double *dp = get_vec(ctx, level, d);
  SellMat *A = get_mat(ctx, level, M);
  if (!vp || !dp || !A) return UGGPU_DESC_MISMATCH;
  if (vp == dp) return uggpu_fail(UGGPU_DESC_MISMATCH, "l_luiter: result and right-hand side are the same vector");
  UG_TRY(tri_solve(ctx, level, M, 0, vp, dp, nullptr, 2));
  UG_TRY(tri_solve(ctx, level, M, 1, vp, vp, nullptr, 3));
  return check_device_error(ctx);
}

// One smoothing step of class `kind` in defect-correction form: x = correction, b updated to the new defect.
//   UGGPU_SM_JAC  Smoother iter.cc:817 + JacobiStep :911     UGGPU_SM_GS   Smoother + GSStep :1039
//   UGGPU_SM_SGS  SGSSmoother :1392 (tmp = NP_SGS_t)          UGGPU_SM_SOR  SORSmoother :4786 + SORStep :4744
//   UGGPU_SM_ILU  Smoother + ILUStep :5478 (tmp = matrix handle of the decomposition)
extern "C" int uggpu_smooth(uggpu_ctx *ctx, int level, int kind, int x, int b, int A, const double *damp, int tmp)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (kind == UGGPU_SM_JAC) return uggpu_jac_smooth(ctx, level, x, b, A, damp);
  double *xp = get_vec(ctx, level, x), *bp = get_vec(ctx, level, b);
  if (!xp || !bp) return UGGPU_DESC_MISMATCH;
  const Damp dm = mkdamp(damp, L->bs);
  switch (kind) {
    case UGGPU_SM_GS:
      UG_TRY(tri_solve(ctx, level, A, 0, xp, bp, nullptr));
      UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, xp, nullptr, dm));
      return k_dmatmul(ctx, level, 2, 0, b, A, x);
    case UGGPU_SM_SOR:
      UG_TRY(tri_solve(ctx, level, A, 0, xp, bp, damp));
      return k_dmatmul(ctx, level, 2, 0, b, A, x);
    case UGGPU_SM_ILU:                                  // Smoother iter.cc:817 + ILUStep :5478; tmp = handle of the decomposed matrix
      if (!get_mat(ctx, level, tmp)) return UGGPU_DESC_MISMATCH;
      UG_TRY(tri_solve(ctx, level, tmp, 0, xp, bp, nullptr, 2));
      UG_TRY(tri_solve(ctx, level, tmp, 1, xp, xp, nullptr, 3));
      UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, xp, nullptr, dm));
      return k_dmatmul(ctx, level, 2, 0, b, A, x);
    case UGGPU_SM_SGS: {
      UG_TRY(uggpu_vec_alloc(ctx, level, tmp));
      double *tp = get_vec(ctx, level, tmp);
      if (!tp) return UGGPU_DESC_MISMATCH;
      if (tp == xp || tp == bp) return uggpu_fail(UGGPU_DESC_MISMATCH, "sgs: the work vector aliases x or b");
      UG_TRY(tri_solve(ctx, level, A, 0, tp, bp, nullptr));
      UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, tp, nullptr, dm));
      UG_TRY(k_dmatmul(ctx, level, 2, 0, b, A, tmp));
      UG_TRY(tri_solve(ctx, level, A, 1, xp, bp, nullptr));
      UG_TRY(k_vec_op(ctx, level, 0, VOP_SCALX, xp, nullptr, dm));
      UG_TRY(k_dmatmul(ctx, level, 2, 0, b, A, x));
      return k_vec_op(ctx, level, 0, VOP_ADD, xp, tp, mkdamp(nullptr, 0));
    }
  }
  return uggpu_fail(UGGPU_ERROR, "unknown smoother class %d", kind);
}
