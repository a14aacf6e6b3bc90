// galerkin.cu -- Galerkin coarse-grid operator: AssembleGalerkinByMatrix np/algebra/transgrid.cc:1575-1700
// (symmetric = 0) after dmatset(coarse, 0), the way `npcheck $G` calls it (np/algebra/npcheck.cc:375-379):
//     A_{l-1} = P^T A_l P     on the interpolation stencils of level l, accumulated in the reference's order.
//
// The reference scatters: fine rows v in list order, their entries m = (v,w) in list order, the interpolation entries im of v, those
// jm of w:  coarse(iv,jv) += (m * im) * jm.  A coarse entry therefore receives its terms ordered by (v, position of m in row v,
// position of jm in the interpolation row of w) -- im is unique per (v, iv).  Here ONE THREAD PER COARSE ROW gathers them in that
// order: the fine rows that interpolate from iv (the transpose of P, built once per call by a stable radix sort of (coarse, fine)
// pairs, so they come ascending), for each its matrix row, for each entry the interpolation row of the neighbour.  No atomics, no
// reduction tree: the same additions in the same order as the sequential loop -- bit-identical values.  Blocks follow :1629-1700:
// per (i,j) the sum over k,l of (IM[i][k] * M[k][l]) * JM[j][l] with the c*I interpolation blocks written out (their zeros take part).
// The product must stay on the coarse pattern (it does on nested geometric hierarchies); the reference would create the missing
// connections, this implementation reports them (UGGPU_ERROR).  A setup operation: rows are gathered uncoalesced.
#include <cub/device/device_radix_sort.cuh>       // before uggpu_internal.h (its SLICE macro)
#include <cub/device/device_scan.cuh>

#include "uggpu_internal.h"

__device__ __forceinline__ double sell_weight(const SellView &T, int r, int j)
{
  const int64_t idx = slice_off(T, r >> 5) + (int64_t)j * 32 + (r & 31);
  return T.vcode ? T.vtable[T.vcode[idx]] : T.val[idx];
}

__global__ void k_gal_len(int n, const uint16_t *__restrict__ rowlen, int64_t *__restrict__ len)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n) len[r] = r < n ? rowlen[r] : 0;
}

// (coarse column, fine row) of every interpolation entry, rows ascending; cnt[c] = entries of coarse vector c
__global__ void k_gal_pairs(SellView P, const int64_t *__restrict__ prp, int32_t *__restrict__ key, int32_t *__restrict__ val, int *__restrict__ cnt)
{
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= P.n) return;
  const int len = P.rowlen[v];
  const ColIter ci = col_iter(P, v);
  for (int j = 0; j < len; j++) {
    const int c = col_at(ci, j);
    key[prp[v] + j] = c; val[prp[v] + j] = v;
    atomicAdd(&cnt[c], 1);
  }
}

__global__ void k_gal_cnt64(int n, const int *__restrict__ cnt, int64_t *__restrict__ out)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= n) out[r] = r < n ? cnt[r] : 0;
}

#define GAL_ACC 32
template <int BS>
__global__ void __launch_bounds__(128) k_galerkin(SellView Ac, double *cval, SellView Af, const double *__restrict__ fval, SellView P,
                                                  const int64_t *__restrict__ tptr, const int32_t *__restrict__ tfine, int *err)
{
  constexpr int BB = BS * BS;
  const int iv = blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= Ac.n) return;
  const int lenc = Ac.rowlen[iv];
  const ColIter cic = col_iter(Ac, iv);
  double *cr = cval + slice_off(Ac, iv >> 5) * BB + (iv & 31);        // component k of entry t of this row: cr[(t*BB + k)*32]
  // Rows of up to GAL_ACC entries accumulate in a thread-private array (the same additions in the same order; one store per entry at the
  // end) and keep their column indices there, too: the first version read-modify-wrote the coarse values in HBM for every one of the
  // ~430 terms of a coarse row (186 ms on the finest level of the 513^3 hierarchy)
  const bool local = lenc <= GAL_ACC;
  double acc[GAL_ACC * BB];
  int ccol[GAL_ACC];
  if (local) {
    for (int t = 0; t < lenc; t++) {
      ccol[t] = col_at(cic, t);
      for (int k = 0; k < BB; k++) acc[t * BB + k] = 0.0;
    }
  } else {
    for (int t = 0; t < lenc; t++)
      for (int k = 0; k < BB; k++) cr[((size_t)t * BB + k) * 32] = 0.0;  // dmatset(level-1, A, 0.0)
  }
  for (int64_t q = tptr[iv]; q < tptr[iv + 1]; q++) {
    const int v = tfine[q];
    // im = the interpolation entry (v, iv)
    const int lenp = P.rowlen[v];
    const ColIter cip = col_iter(P, v);
    double wim = 0.0;
    for (int j = 0; j < lenp; j++) if (col_at(cip, j) == iv) { wim = sell_weight(P, v, j); break; }
    const int lenf = Af.rowlen[v];
    const ColIter cif = col_iter(Af, v);
    const double *fr = fval + slice_off(Af, v >> 5) * BB + (v & 31);
    for (int e = 0; e < lenf; e++) {
      const int w = col_at(cif, e);
      double M[BB];
#pragma unroll
      for (int k = 0; k < BB; k++) M[k] = fr[((size_t)e * BB + k) * 32];
      const int lenw = P.rowlen[w];
      const ColIter ciw = col_iter(P, w);
      for (int j = 0; j < lenw; j++) {
        const int jv = col_at(ciw, j);
        const double wjm = sell_weight(P, w, j);
        int t = -1;                                                     // GetMatrix(iv, jv)
        if (local) { for (int u = 0; u < lenc; u++) if (ccol[u] == jv) { t = u; break; } }
        else { for (int u = 0; u < lenc; u++) if (col_at(cic, u) == jv) { t = u; break; } }
        if (t < 0) { atomicExch(err, UGGPU_ERROR); continue; }          // the reference would create the connection
        if (BS == 1) {
          const double fac = M[0] * wim;
          const double p = fac * wjm;
          if (local) acc[t] = acc[t] + p;
          else cr[(size_t)t * 32] = cr[(size_t)t * 32] + p;
        } else {
#pragma unroll
          for (int i = 0; i < BS; i++)
#pragma unroll
            for (int jj = 0; jj < BS; jj++) {
              double sum = 0.0;
#pragma unroll
              for (int k = 0; k < BS; k++)
#pragma unroll
                for (int l = 0; l < BS; l++) {
                  const double a = (k == i ? wim : 0.0) * M[k * BS + l];
                  const double p = a * (l == jj ? wjm : 0.0);
                  sum += p;
                }
              if (local) acc[t * BB + i * BS + jj] = acc[t * BB + i * BS + jj] + sum;
              else cr[((size_t)t * BB + i * BS + jj) * 32] = cr[((size_t)t * BB + i * BS + jj) * 32] + sum;
            }
        }
      }
    }
  }
  if (local)
    for (int t = 0; t < lenc; t++)
      for (int k = 0; k < BB; k++) cr[((size_t)t * BB + k) * 32] = acc[t * BB + k];
}

// ---- pattern growth (CreateExtraConnection, transgrid.cc:1615-1617 / :1649-1651) ------------------------------------------------------
// The SYMBOLIC half of the product runs on the host: which connections the reference creates, and where they end up in the rows' lists,
// is decided by the order of a sequential traversal (every new connection goes to the second place of both rows, gm/algebra.cc:1051-1078).
// A row ends up as: diagonal, the connections created by the product in reverse order of creation, its old off-diagonal entries.  The
// NUMERIC half is the kernel above on the new pattern.  A setup step of AMG hierarchies (np/procs/amgtransfer.cc:915-925), sizes of a few
// 10^5 rows; a device form (first-touch times of the pairs by a segmented min + sort) is the next step, DESIGN.md 9.
#include <vector>

namespace {
// set of directed pairs (row << 32 | column): open addressing, linear probing, grows at 1/2 load.  (std::unordered_set costs ~400 ns per
// term of the product here; this table ~30 ns.)
struct PairSet {
  std::vector<uint64_t> tab;
  size_t mask = 0, count = 0;
  explicit PairSet(size_t expect) { size_t cap = 1024; while (cap < 2 * expect + 16) cap <<= 1; tab.assign(cap, 0); mask = cap - 1; }
  static uint64_t mix(uint64_t k) { k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33; return k; }
  void grow()
  {
    std::vector<uint64_t> old;
    old.swap(tab);
    tab.assign(old.size() * 2, 0); mask = tab.size() - 1; count = 0;
    for (uint64_t k : old) if (k) insert_key(k);
  }
  bool insert_key(uint64_t k)        // k != 0
  {
    size_t i = mix(k) & mask;
    while (tab[i]) { if (tab[i] == k) return false; i = (i + 1) & mask; }
    tab[i] = k;
    if (++count * 2 > tab.size()) grow();
    return true;
  }
  bool insert(int32_t r, int32_t c) { return insert_key((((uint64_t)(uint32_t)r << 32) | (uint32_t)c) + 1); }
};
}  // namespace

static int galerkin_pattern_impl(int nf, int nc, const int32_t *a_rowptr, const int32_t *a_col, const int32_t *p_rowptr, const int32_t *p_col,
                                 const int32_t *start_rowptr, const int32_t *start_col, int32_t *out_rowptr, int32_t *out_col, std::vector<int32_t> *out_vec);

extern "C" int uggpu_galerkin_pattern(int nf, int nc, const int32_t *a_rowptr, const int32_t *a_col, const int32_t *p_rowptr, const int32_t *p_col,
                                      const int32_t *start_rowptr, const int32_t *start_col, int32_t *out_rowptr, int32_t *out_col)
{
  return galerkin_pattern_impl(nf, nc, a_rowptr, a_col, p_rowptr, p_col, start_rowptr, start_col, out_rowptr, out_col, nullptr);
}

// out_vec (internal callers): receives the columns, sized here -- one traversal instead of the two of the count-then-fill protocol
static int galerkin_pattern_impl(int nf, int nc, const int32_t *a_rowptr, const int32_t *a_col, const int32_t *p_rowptr, const int32_t *p_col,
                                 const int32_t *start_rowptr, const int32_t *start_col, int32_t *out_rowptr, int32_t *out_col, std::vector<int32_t> *out_vec)
{
  if (nf < 0 || nc < 0 || !a_rowptr || !a_col || !p_rowptr || !p_col || !out_rowptr) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: null argument");
  if ((start_rowptr == nullptr) != (start_col == nullptr)) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: start pattern needs both arrays");
  PairSet have(start_rowptr ? (size_t)start_rowptr[nc] * 2 : (size_t)nc * 16);      // directed pairs present so far
  std::vector<std::vector<int32_t> > created((size_t)nc);  // per row: the columns of its new connections in creation order
  if (start_rowptr) {
    for (int i = 0; i < nc; i++) {
      if (start_rowptr[i + 1] <= start_rowptr[i] || start_col[start_rowptr[i]] != i) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: row %d does not start with its diagonal entry", i);
      for (int e = start_rowptr[i]; e < start_rowptr[i + 1]; e++) {
        if (start_col[e] < 0 || start_col[e] >= nc) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: column out of range in row %d", i);
        have.insert(i, start_col[e]);
      }
    }
  } else {
    for (int i = 0; i < nc; i++) have.insert(i, i);
  }
  for (int v = 0; v < nf; v++)
    for (int e = a_rowptr[v]; e < a_rowptr[v + 1]; e++) {
      const int w = a_col[e];
      if (w < 0 || w >= nf) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: fine column out of range in row %d", v);
      for (int ie = p_rowptr[v]; ie < p_rowptr[v + 1]; ie++) {
        const int32_t iv = p_col[ie];
        if (iv < 0 || iv >= nc) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: interpolation column out of range in row %d", v);
        for (int je = p_rowptr[w]; je < p_rowptr[w + 1]; je++) {
          const int32_t jv = p_col[je];
          if (jv < 0 || jv >= nc) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: interpolation column out of range in row %d", w);
          if (iv == jv || !have.insert(iv, jv)) continue;             // GetMatrix(iv, jv) finds it (the diagonal always exists)
          have.insert(jv, iv);                                       // one CONNECTION holds both directions
          created[iv].push_back(jv);
          created[jv].push_back(iv);
        }
      }
    }
  int64_t total = 0;
  out_rowptr[0] = 0;
  for (int i = 0; i < nc; i++) {
    total += (start_rowptr ? start_rowptr[i + 1] - start_rowptr[i] : 1) + (int64_t)created[i].size();
    if (total > 2147483647LL) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin_pattern: more than 2^31 - 1 entries");
    out_rowptr[i + 1] = (int32_t)total;
  }
  if (out_vec) { out_vec->resize((size_t)total + 1); out_col = out_vec->data(); }
  if (!out_col) return 0;
  for (int i = 0; i < nc; i++) {
    int32_t *o = out_col + out_rowptr[i];
    *o++ = i;
    for (size_t k = created[i].size(); k-- > 0;) *o++ = created[i][k];
    if (start_rowptr) for (int e = start_rowptr[i] + 1; e < start_rowptr[i + 1]; e++) *o++ = start_col[e];
  }
  return 0;
}

// gives matrix A of level-1 the pattern the product needs: from its present pattern, or -- no such matrix yet -- from one diagonal entry per row
static int galerkin_grow(uggpu_ctx *ctx, int level, int A)
{
  Level *L = get_level(ctx, level), *C = get_level(ctx, level - 1);
  SellMat *Af = get_mat(ctx, level, A), *Ac = get_mat_quiet(ctx, level - 1, A);
  if (!L || !C || !Af) return UGGPU_DESC_MISMATCH;
  const int nf = L->n, nc = C->n;
  std::vector<int32_t> arp((size_t)nf + 1), acol((size_t)Af->nnz + 1), prp((size_t)nf + 1), pcol((size_t)L->P.nnz + 1), srp, scol, orp((size_t)nc + 1), ocol;
  UG_TRY(sell_to_host_csr(ctx, Af, arp.data(), acol.data(), nullptr));
  UG_TRY(sell_to_host_csr(ctx, &L->P, prp.data(), pcol.data(), nullptr));
  if (Ac) {
    srp.resize((size_t)nc + 1); scol.resize((size_t)Ac->nnz + 1);
    UG_TRY(sell_to_host_csr(ctx, Ac, srp.data(), scol.data(), nullptr));
  }
  UG_TRY(galerkin_pattern_impl(nf, nc, arp.data(), acol.data(), prp.data(), pcol.data(), Ac ? srp.data() : nullptr, Ac ? scol.data() : nullptr, orp.data(), nullptr, &ocol));
  return uggpu_mat_set_pattern(ctx, level - 1, A, orp.data(), ocol.data());      // values zero: the product writes all of them
}

static int galerkin_product(uggpu_ctx *ctx, int level, int A, bool *pattern_miss);

extern "C" int uggpu_galerkin(uggpu_ctx *ctx, int level, int A)
{
  Level *L = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!L || !C) return UGGPU_NO_COARSER_GRID;
  if (!get_mat(ctx, level, A)) return UGGPU_DESC_MISMATCH;
  if (!L->P.valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "level %d has no interpolation stencil", level);
  if (ctx->comm && (L->partitioned || C->partitioned)) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin runs on one GPU (level %d is partitioned)", level);
  if (L->n == 0 || C->n == 0) return 0;
  if (L->bs < 1 || L->bs > 3 || C->bs != L->bs) return uggpu_fail(UGGPU_BLOCK_TOO_LARGE, "uggpu_galerkin: block size %d", L->bs);
  // a coarse level without this matrix (a fresh algebraic level, amgtransfer.cc:915): the pattern comes from the product alone
  if (!get_mat_quiet(ctx, level - 1, A)) UG_TRY(galerkin_grow(ctx, level, A));
  bool miss = false;
  UG_TRY(galerkin_product(ctx, level, A, &miss));
  if (miss) {          // the product leaves the pattern: create the connections like the reference and run it again
    UG_TRY(galerkin_grow(ctx, level, A));
    UG_TRY(galerkin_product(ctx, level, A, &miss));
    if (miss) return uggpu_fail(UGGPU_ERROR, "uggpu_galerkin: the grown pattern of level %d still misses a connection", level - 1);
  }
  return 0;
}

static int galerkin_product(uggpu_ctx *ctx, int level, int A, bool *pattern_miss)
{
  Level *L = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!L || !C) return UGGPU_NO_COARSER_GRID;
  SellMat *Af = get_mat(ctx, level, A), *Ac = get_mat(ctx, level - 1, A);
  if (!Af || !Ac) return UGGPU_DESC_MISMATCH;
  *pattern_miss = false;
  cudaStream_t st = ctx->stream;
  const int nf = L->n, nc = C->n;
  const int64_t zp = L->P.nnz;
  int64_t *len = nullptr, *prp = nullptr, *tptr = nullptr;
  int32_t *key = nullptr, *val = nullptr, *key2 = nullptr, *tfine = nullptr;
  int *cnt = nullptr;
  void *tmp = nullptr; size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
  const size_t nmax = (size_t)(nf > nc ? nf : nc) + 1, zz = (size_t)(zp > 0 ? zp : 1);
  int rc = 0;
#define GT(expr) do { if (!rc) rc = (expr); } while (0)
#define GC(expr) do { if (!rc) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); } } while (0)
  GT(dalloc(ctx, &len, nmax)); GT(dalloc(ctx, &prp, nmax)); GT(dalloc(ctx, &tptr, (size_t)nc + 1));
  GT(dalloc(ctx, &key, zz)); GT(dalloc(ctx, &val, zz)); GT(dalloc(ctx, &key2, zz)); GT(dalloc(ctx, &tfine, zz)); GT(dalloc(ctx, &cnt, (size_t)nc));
  if (!rc) {
    int bits = 1;
    while ((1ll << bits) < (long long)nc && bits < 31) bits++;
    GC(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len, prp, nf + 1, st));
    GC(cub::DeviceScan::ExclusiveSum(nullptr, tb2, len, tptr, nc + 1, st));
    GC(cub::DeviceRadixSort::SortPairs(nullptr, tb3, key, key2, val, tfine, (int)zp, 0, bits, st));
    if (tb2 > tmp_bytes) tmp_bytes = tb2;
    if (tb3 > tmp_bytes) tmp_bytes = tb3;
    GT(dev_alloc(ctx, &tmp, tmp_bytes ? tmp_bytes : 1));
    // rows of P -> offsets; (coarse, fine) pairs in row order; a stable sort by the coarse index leaves the fine rows ascending
    if (!rc) { k_gal_len<<<(nf + 256) / 256, 256, 0, st>>>(nf, L->P.rowlen, len); ctx->launches++; }
    GC(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, prp, nf + 1, st));
    GC(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)nc, st));
    if (!rc) { k_gal_pairs<<<(nf + 255) / 256, 256, 0, st>>>(view(L->P), prp, key, val, cnt); ctx->launches++; }
    if (zp > 0) GC(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, val, tfine, (int)zp, 0, bits, st));
    if (!rc) { k_gal_cnt64<<<(nc + 256) / 256, 256, 0, st>>>(nc, cnt, len); ctx->launches++; }
    GC(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, tptr, nc + 1, st));
    // the coarse values are about to change: nothing derived from them may survive
    GT(sell_drop_shared_values(ctx, Ac));
    GT(sell_free_schedules(ctx, Ac));
    if (!rc) {
      const int blocks = (nc + 127) / 128;
      const SellView Acv = view(*Ac), Afv = view(*Af), Pv = view(L->P);
      ProfScope ps(ctx, UGGPU_K_GALERKIN, level, 12.0 * (double)Af->nnz * Af->bb + 24.0 * (double)zp + 12.0 * (double)Ac->nnz * Ac->bb + 4.0 * ((double)nf + nc));
      switch (L->bs) {
        case 1: k_galerkin<1><<<blocks, 128, 0, st>>>(Acv, Ac->val, Afv, Af->val, Pv, tptr, tfine, ctx->derr); break;
        case 2: k_galerkin<2><<<blocks, 128, 0, st>>>(Acv, Ac->val, Afv, Af->val, Pv, tptr, tfine, ctx->derr); break;
        default: k_galerkin<3><<<blocks, 128, 0, st>>>(Acv, Ac->val, Afv, Af->val, Pv, tptr, tfine, ctx->derr); break;
      }
      ctx->launches++;
      GC(cudaGetLastError());
    }
    bool launched = false;
    if (!rc) {
      launched = true;
      rc = check_device_error(ctx);
      if (rc == UGGPU_ERROR) { *pattern_miss = true; rc = 0; }      // a term without a coarse entry (the kernel's only report): the caller grows the pattern
    }
    // the coarse values have changed whatever the kernel reported: diagonal array and value generation follow them, also after a failed call
    if (launched) {
      const int rc2 = sell_update_diag(ctx, Ac);
      if (!rc) rc = rc2;
      if (!rc) rc = sell_share_values(ctx, Ac);
    }
    cudaStreamSynchronize(st);
  }
#undef GT
#undef GC
  if (tmp) dev_free(ctx, tmp, tmp_bytes ? tmp_bytes : 1);
  if (len) dfree(ctx, len, nmax);
  if (prp) dfree(ctx, prp, nmax);
  if (tptr) dfree(ctx, tptr, (size_t)nc + 1);
  if (key) dfree(ctx, key, zz);
  if (val) dfree(ctx, val, zz);
  if (key2) dfree(ctx, key2, zz);
  if (tfine) dfree(ctx, tfine, zz);
  if (cnt) dfree(ctx, cnt, (size_t)nc);
  return rc;
}
