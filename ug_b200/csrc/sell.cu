// sell.cu -- CSR/BSR (canonical, host-facing) <-> SELL-32 (device-resident) conversion.
//
// The canonical layout of SURVEY.md 8(a') -- rows in FIRSTVECTOR->SUCCVC order, entries in VSTART->MNEXT
// order -- is what crosses the C-ABI (uggpu_mat_set / uggpu_mat_get round-trip bit-exactly).  On the device the
// same entries are stored slice-interleaved so that thread-per-row kernels are coalesced (uggpu_internal.h).
#include "uggpu_internal.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

void sell_layout(const std::vector<int> &width, std::vector<int64_t> &sp, int *maxlen, int *fixed_w)
{
  const size_t nsl = width.size();
  int mx = 0;
  int64_t padded = 0;
  for (size_t s = 0; s < nsl; s++) { if (width[s] > mx) mx = width[s]; padded += (int64_t)width[s] * 32; }
  const int64_t fixed = (int64_t)nsl * 32 * mx;
  const bool fix = mx > 0 && fixed <= padded + padded / 8 && !getenv("UGGPU_NO_FIXED_WIDTH");
  sp.assign(nsl + 1, 0);
  for (size_t s = 0; s < nsl; s++) sp[s + 1] = sp[s] + (int64_t)(fix ? mx : width[s]) * 32;
  *maxlen = mx;
  *fixed_w = fix ? mx : 0;
}

__global__ void k_sell_rowlen(int n, const int64_t *__restrict__ rowptr, uint16_t *__restrict__ rowlen, int *__restrict__ width, int *err)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int lane = threadIdx.x & 31;
  int64_t len = 0;
  if (r < n) {
    len = rowptr[r + 1] - rowptr[r];
    if (len < 0 || len > 65535) { atomicExch(err, UGGPU_ERROR); len = 0; }
    rowlen[r] = (uint16_t)len;
  }
  int w = (int)len;
  for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
  if (lane == 0 && (r >> 5) < (n + 31) / 32) width[r >> 5] = w;
}

// one thread per row; padding entries get col = own row (clamped) and value 0
__global__ void k_sell_fill(int n, int bb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ ccol, const double *__restrict__ cval,
                            const int64_t *__restrict__ slice_ptr, int32_t *__restrict__ col, double *__restrict__ val)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int s = r >> 5, lane = r & 31;
  int nsl = (n + 31) >> 5;
  if (s >= nsl) return;
  int64_t sp = slice_ptr[s];
  int w = (int)((slice_ptr[s + 1] - sp) >> 5);
  int64_t rp = 0;
  int len = 0;
  if (r < n) { rp = rowptr[r]; len = (int)(rowptr[r + 1] - rp); }
  int padcol = r < n ? r : 0;
  for (int j = 0; j < w; j++) {
    int64_t dst = sp + (int64_t)j * 32;
    if (j < len) {
      col[dst + lane] = ccol[rp + j];
      for (int k = 0; k < bb; k++) val[dst * bb + (int64_t)k * 32 + lane] = cval[(rp + j) * bb + k];
    } else {
      col[dst + lane] = padcol;
      for (int k = 0; k < bb; k++) val[dst * bb + (int64_t)k * 32 + lane] = 0.0;
    }
  }
}

__global__ void k_sell_set_values(int n, int bb, const int64_t *__restrict__ rowptr, const double *__restrict__ cval,
                                  const int64_t *__restrict__ slice_ptr, double *__restrict__ val)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int s = r >> 5, lane = r & 31;
  int64_t sp = slice_ptr[s];
  int64_t rp = rowptr[r];
  int len = (int)(rowptr[r + 1] - rp);
  for (int j = 0; j < len; j++) {
    int64_t dst = sp + (int64_t)j * 32;
    for (int k = 0; k < bb; k++) val[dst * bb + (int64_t)k * 32 + lane] = cval[(rp + j) * bb + k];
  }
}

__global__ void k_sell_to_csr(SellView A, int bb, const int64_t *__restrict__ rowptr, int32_t *__restrict__ ccol, double *__restrict__ cval)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  int s = r >> 5, lane = r & 31;
  int64_t sp = A.slice_ptr[s];
  int64_t rp = rowptr[r];
  int len = (int)(rowptr[r + 1] - rp);
  const ColIter ci = col_iter(A, r);
  for (int j = 0; j < len; j++) {
    int64_t src = sp + (int64_t)j * 32;
    ccol[rp + j] = col_at(ci, j);
    for (int k = 0; k < bb; k++) cval[(rp + j) * bb + k] = A.val[src * bb + (int64_t)k * 32 + lane];
  }
}

// ---- column-index compression ------------------------------------------------------------------------------------------
// One warp per slice.  flag[s] = 1 when, for every slice column j, all rows that have an entry j hold the same
// distance col - row (and the slice is at most 32 columns wide, so that a warp can keep the distances in one register
// per lane); cnt[s] = true entries of the slice; hash[s] = 64-bit hash of the slice's distance vector.
__global__ void k_sell_uniform_flag(int n, const int64_t *__restrict__ slice_ptr, const uint16_t *__restrict__ rowlen, const int32_t *__restrict__ col,
                                    uint8_t *__restrict__ flag, int *__restrict__ cnt, unsigned long long *__restrict__ hash)
{
  const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= (n + 31) / 32) return;
  const int r = s * 32 + lane;
  const int len = r < n ? rowlen[r] : 0;
  const int64_t sp = slice_ptr[s];
  const int w = (int)((slice_ptr[s + 1] - sp) >> 5);
  bool ok = true;
  unsigned long long h = 1469598103934665603ull ^ (unsigned long long)w;
  for (int j = 0; j < w; j++) {
    const bool has = j < len;
    const int d = has ? col[sp + (int64_t)j * 32 + lane] - r : 0;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    const int ref = __shfl_sync(0xffffffffu, d, m ? __ffs(m) - 1 : 0);
    if (has && d != ref) ok = false;
    h = (h ^ (unsigned long long)(unsigned)ref) * 1099511628211ull;
    h ^= h >> 29;
  }
  ok = __all_sync(0xffffffffu, ok);
  int c = len;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) { flag[s] = (ok && w > 0 && w <= 32) ? 1 : 0; cnt[s] = c; hash[s] = h; }
}

// writes the new column array; uniform slices that share a table write identical words to the same place
__global__ void k_sell_compact_cols(int n, const int64_t *__restrict__ slice_ptr, const uint16_t *__restrict__ rowlen, const int32_t *__restrict__ col,
                                    const int64_t *__restrict__ col_ptr, int32_t *__restrict__ ncol)
{
  const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= (n + 31) / 32) return;
  const int r = s * 32 + lane;
  const int len = r < n ? rowlen[r] : 0;
  const int64_t sp = slice_ptr[s];
  const int w = (int)((slice_ptr[s + 1] - sp) >> 5);
  const int64_t cp = col_ptr[s];
  if (cp >= 0) {
    for (int j = 0; j < w; j++) ncol[cp + (int64_t)j * 32 + lane] = col[sp + (int64_t)j * 32 + lane];
  } else {
    for (int j = 0; j < w; j++) {
      const bool has = j < len;
      const int d = has ? col[sp + (int64_t)j * 32 + lane] - r : 0;
      const unsigned m = __ballot_sync(0xffffffffu, has);
      const int ref = __shfl_sync(0xffffffffu, d, m ? __ffs(m) - 1 : 0);
      if (lane == 0) ncol[~cp + j] = ref;
    }
  }
}

// decodes every true entry from the new arrays and compares it with the explicit original
__global__ void k_sell_verify_cols(int n, const int64_t *__restrict__ slice_ptr, const uint16_t *__restrict__ rowlen, const int32_t *__restrict__ col,
                                   const int64_t *__restrict__ col_ptr, const int32_t *__restrict__ ncol, unsigned long long *__restrict__ bad)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int s = r >> 5, lane = r & 31;
  const int64_t sp = slice_ptr[s], cp = col_ptr[s];
  const int len = rowlen[r];
  int wrong = 0;
  for (int j = 0; j < len; j++) {
    const int c = cp < 0 ? ncol[~cp + j] + r : ncol[cp + (int64_t)j * 32 + lane];
    if (c != col[sp + (int64_t)j * 32 + lane]) wrong++;
  }
  if (wrong) atomicAdd(bad, (unsigned long long)wrong);
}

int sell_compress_cols(uggpu_ctx *ctx, SellMat *m)
{
  if (m->n <= 0 || m->col_ptr != m->slice_ptr || m->padded == 0) return 0;
  if (getenv("UGGPU_NO_COL_COMPRESSION")) return 0;
  cudaStream_t st = ctx->stream;
  const size_t nsl = (size_t)(m->n + 31) / 32;
  uint8_t *d_flag = nullptr; int *d_cnt = nullptr; unsigned long long *d_hash = nullptr;
  UG_TRY(dalloc(ctx, &d_flag, nsl));
  UG_TRY(dalloc(ctx, &d_cnt, nsl));
  UG_TRY(dalloc(ctx, &d_hash, nsl + 1));       // [nsl]: mismatch counter of the verification
  const int blocks = (int)((nsl * 32 + 255) / 256);
  k_sell_uniform_flag<<<blocks, 256, 0, st>>>(m->n, m->slice_ptr, m->rowlen, m->col, d_flag, d_cnt, d_hash);
  KCHECK(ctx);
  std::vector<uint8_t> flag(nsl);
  std::vector<int> cnt(nsl);
  std::vector<unsigned long long> hash(nsl);
  std::vector<int64_t> sp(nsl + 1), cp(nsl);
  CUDA_TRY(cudaMemcpyAsync(flag.data(), d_flag, nsl, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, nsl * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(hash.data(), d_hash, nsl * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(sp.data(), m->slice_ptr, (nsl + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  UG_TRY(dfree(ctx, d_flag, nsl));
  UG_TRY(dfree(ctx, d_cnt, nsl));
  int rc = 0;
  // slices with equal distance vectors share one table (found by hash, confirmed by the verification pass below; a
  // hash collision makes the pass fail and the tables are then stored per slice)
  for (int dedupe = 1; dedupe >= 0; dedupe--) {
    std::map<unsigned long long, int64_t> tables;
    int64_t uoff = 0, words = 0, uni = 0;
    // tables first, explicit slices behind them (128-byte aligned)
    for (size_t s = 0; s < nsl; s++) {
      if (!flag[s]) continue;
      const int64_t w = (sp[s + 1] - sp[s]) >> 5;
      uni++;
      if (dedupe) {
        auto it = tables.find(hash[s]);
        if (it != tables.end()) { cp[s] = ~it->second; continue; }
        tables[hash[s]] = uoff;
      }
      cp[s] = ~uoff; uoff += (w + 3) & ~(int64_t)3; words += w;      // 16-byte granules; a shared table is read from HBM once
    }
    int64_t off = (uoff + 31) & ~(int64_t)31;
    for (size_t s = 0; s < nsl; s++) {
      if (flag[s]) continue;
      const int64_t w = (sp[s + 1] - sp[s]) >> 5;
      cp[s] = off; off += w * 32; words += cnt[s];
    }
    off += 32;                                   // a warp reads a table as 32 words
    if (uni == 0 || off >= m->padded) break;
    int64_t *d_cp = nullptr; int32_t *ncol = nullptr;
    if ((rc = dalloc(ctx, &d_cp, nsl)) != 0) break;
    if ((rc = dalloc(ctx, &ncol, (size_t)off)) != 0) { dfree(ctx, d_cp, nsl); break; }
    cudaMemsetAsync(ncol, 0, (size_t)off * sizeof(int32_t), st);
    cudaMemsetAsync(d_hash + nsl, 0, sizeof(unsigned long long), st);
    cudaMemcpyAsync(d_cp, cp.data(), nsl * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    k_sell_compact_cols<<<blocks, 256, 0, st>>>(m->n, m->slice_ptr, m->rowlen, m->col, d_cp, ncol);
    ctx->launches++;
    k_sell_verify_cols<<<(m->n + 255) / 256, 256, 0, st>>>(m->n, m->slice_ptr, m->rowlen, m->col, d_cp, ncol, d_hash + nsl);
    ctx->launches++;
    unsigned long long bad = 1;
    cudaMemcpyAsync(&bad, d_hash + nsl, sizeof bad, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = uggpu_fail(UGGPU_CUDA_ERROR, "column compression: %s", cudaGetErrorString(e)); dfree(ctx, d_cp, nsl); dfree(ctx, ncol, (size_t)off); break; }
    if (bad) { dfree(ctx, d_cp, nsl); dfree(ctx, ncol, (size_t)off); continue; }     // hash collision: retry without sharing
    dfree(ctx, m->col, (size_t)m->col_len);
    m->col = ncol; m->col_len = off; m->col_ptr = d_cp; m->col_words = words; m->uniform_slices = uni;
    break;
  }
  dfree(ctx, d_hash, nsl + 1);
  return rc;
}

// ---- shared value tables of uniform slices (matrices) ------------------------------------------------------------------------
// One warp per slice with uniform column distances.  vflag[s] = 1 when, for every slice column j and block component k, all rows
// that have an entry j hold the same bit pattern; cnt[s] = true entries of the slice; hash[s] = 64-bit hash of the value vector.
__global__ void k_sell_vuniform_flag(int n, int bb, const int64_t *__restrict__ slice_ptr, const int64_t *__restrict__ col_ptr, const uint16_t *__restrict__ rowlen,
                                     const double *__restrict__ val, uint8_t *__restrict__ vflag, int *__restrict__ cnt, unsigned long long *__restrict__ hash)
{
  const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= (n + 31) / 32) return;
  const int r = s * 32 + lane;
  const int len = r < n ? rowlen[r] : 0;
  const int64_t sp = slice_ptr[s];
  const int w = (int)((slice_ptr[s + 1] - sp) >> 5);
  int c = len;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  bool ok = col_ptr[s] < 0;
  unsigned long long h = 1469598103934665603ull ^ (unsigned long long)(w * 16 + bb);
  if (ok)
    for (int j = 0; j < w; j++) {
      const bool has = j < len;
      const unsigned m = __ballot_sync(0xffffffffu, has);
      const int src = m ? __ffs(m) - 1 : 0;
      for (int k = 0; k < bb; k++) {
        const unsigned long long bits = has ? (unsigned long long)__double_as_longlong(val[(sp + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane]) : 0ull;
        const unsigned long long ref = __shfl_sync(0xffffffffu, bits, src);
        if (has && bits != ref) ok = false;
        h = (h ^ ref) * 1099511628211ull;
        h ^= h >> 29;
      }
    }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) { vflag[s] = ok && w > 0 ? 1 : 0; cnt[s] = c; hash[s] = h; }
}

// slices that share a table write identical doubles to the same place
__global__ void k_sell_fill_vt(int n, int bb, const int64_t *__restrict__ slice_ptr, const int64_t *__restrict__ col_ptr, const uint16_t *__restrict__ rowlen,
                               const double *__restrict__ val, double *__restrict__ vt)
{
  const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (s >= (n + 31) / 32) return;
  const int64_t cp = col_ptr[s];
  if (cp >= 0 || UG_VALTAB(cp) < 0) return;
  const int64_t vo = UG_VALTAB(cp);
  const int r = s * 32 + lane;
  const int len = r < n ? rowlen[r] : 0;
  const int64_t sp = slice_ptr[s];
  const int w = (int)((slice_ptr[s + 1] - sp) >> 5);
  for (int j = 0; j < w; j++) {
    const bool has = j < len;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    const int src = m ? __ffs(m) - 1 : 0;
    if (lane == src && has)
      for (int k = 0; k < bb; k++) vt[vo + (int64_t)j * bb + k] = val[(sp + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane];
  }
}

// decodes every true entry of the slices with shared values from the tables and compares the bits with the explicit values
__global__ void k_sell_verify_vt(int n, int bb, const int64_t *__restrict__ slice_ptr, const int64_t *__restrict__ col_ptr, const uint16_t *__restrict__ rowlen,
                                 const double *__restrict__ val, const double *__restrict__ vt, unsigned long long *__restrict__ bad)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int s = r >> 5, lane = r & 31;
  const int64_t cp = col_ptr[s];
  if (cp >= 0 || UG_VALTAB(cp) < 0) return;
  const int64_t vo = UG_VALTAB(cp), sp = slice_ptr[s];
  const int len = rowlen[r];
  int wrong = 0;
  for (int j = 0; j < len; j++)
    for (int k = 0; k < bb; k++)
      if (__double_as_longlong(vt[vo + (int64_t)j * bb + k]) != __double_as_longlong(val[(sp + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane])) wrong++;
  if (wrong) atomicAdd(bad, (unsigned long long)wrong);
}

__global__ void k_sell_strip_vt(int nsl, int64_t *__restrict__ col_ptr)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nsl && col_ptr[s] < 0) col_ptr[s] = ~UG_COLTAB(col_ptr[s]);
}

int sell_drop_shared_values(uggpu_ctx *ctx, SellMat *m)
{
  if (m->n <= 0 || m->col_ptr == m->slice_ptr || !m->col_ptr) return 0;
  const int nsl = (m->n + 31) / 32;
  k_sell_strip_vt<<<(nsl + 255) / 256, 256, 0, ctx->stream>>>(nsl, m->col_ptr);
  KCHECK(ctx);
  if (m->vt) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); UG_TRY(dfree(ctx, m->vt, (size_t)m->vt_len)); }
  m->vt_len = 0; m->vshared_slices = 0; m->val_entries = -1;
  m->sten.w = 0; m->sten_slices = 0;
  stx_free(ctx, m);
  delete m->sten3; m->sten3 = nullptr;
  return 0;
}

int sell_share_values(uggpu_ctx *ctx, SellMat *m)
{
  if (m->n <= 0 || m->col_ptr == m->slice_ptr || m->uniform_slices == 0 || m->vcode) return 0;
  UG_TRY(sell_drop_shared_values(ctx, m));
  if (getenv("UGGPU_NO_SHARED_VALUES")) return 0;
  cudaStream_t st = ctx->stream;
  const size_t nsl = (size_t)(m->n + 31) / 32;
  uint8_t *d_flag = nullptr; int *d_cnt = nullptr; unsigned long long *d_hash = nullptr;
  UG_TRY(dalloc(ctx, &d_flag, nsl));
  UG_TRY(dalloc(ctx, &d_cnt, nsl));
  UG_TRY(dalloc(ctx, &d_hash, nsl + 1));       // [nsl]: mismatch counter of the verification
  const int blocks = (int)((nsl * 32 + 255) / 256);
  k_sell_vuniform_flag<<<blocks, 256, 0, st>>>(m->n, m->bb, m->slice_ptr, m->col_ptr, m->rowlen, m->val, d_flag, d_cnt, d_hash);
  KCHECK(ctx);
  std::vector<uint8_t> flag(nsl);
  std::vector<int> cnt(nsl);
  std::vector<unsigned long long> hash(nsl);
  std::vector<int64_t> sp(nsl + 1), cp(nsl), ncp(nsl);
  CUDA_TRY(cudaMemcpyAsync(flag.data(), d_flag, nsl, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, nsl * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(hash.data(), d_hash, nsl * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(sp.data(), m->slice_ptr, (nsl + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cp.data(), m->col_ptr, nsl * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  UG_TRY(dfree(ctx, d_flag, nsl));
  UG_TRY(dfree(ctx, d_cnt, nsl));
  int rc = 0;
  for (int dedupe = 1; dedupe >= 0; dedupe--) {
    std::map<unsigned long long, int64_t> tables;
    int64_t voff = 0, entries = 0, shared = 0;
    for (size_t s = 0; s < nsl; s++) {
      ncp[s] = cp[s];
      if (!flag[s]) { entries += cnt[s]; continue; }
      const int64_t w = (sp[s + 1] - sp[s]) >> 5;
      shared++;
      int64_t vo;
      auto it = dedupe ? tables.find(hash[s]) : tables.end();
      if (it != tables.end()) vo = it->second;
      else {
        vo = voff; voff += (w * m->bb + 1) & ~(int64_t)1; entries += w;      // 16-byte granules; a shared table crosses HBM once
        if (dedupe) tables[hash[s]] = vo;
      }
      ncp[s] = ~(UG_COLTAB(cp[s]) | ((vo + 1) << 32));
    }
    if (shared == 0 || voff >= ((int64_t)1 << 30)) break;
    // worth it only if most of the value stream disappears
    if (entries * 4 > m->nnz * 3) break;
    double *vt = nullptr; int64_t *d_cp = nullptr;
    if ((rc = dalloc(ctx, &vt, (size_t)voff)) != 0) break;
    if ((rc = dalloc(ctx, &d_cp, nsl)) != 0) { dfree(ctx, vt, (size_t)voff); break; }
    cudaMemsetAsync(vt, 0, (size_t)voff * sizeof(double), st);
    cudaMemsetAsync(d_hash + nsl, 0, sizeof(unsigned long long), st);
    cudaMemcpyAsync(d_cp, ncp.data(), nsl * sizeof(int64_t), cudaMemcpyHostToDevice, st);
    k_sell_fill_vt<<<blocks, 256, 0, st>>>(m->n, m->bb, m->slice_ptr, d_cp, m->rowlen, m->val, vt);
    ctx->launches++;
    k_sell_verify_vt<<<(m->n + 255) / 256, 256, 0, st>>>(m->n, m->bb, m->slice_ptr, d_cp, m->rowlen, m->val, vt, d_hash + nsl);
    ctx->launches++;
    unsigned long long bad = 1;
    cudaMemcpyAsync(&bad, d_hash + nsl, sizeof bad, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { rc = uggpu_fail(UGGPU_CUDA_ERROR, "shared value tables: %s", cudaGetErrorString(e)); dfree(ctx, vt, (size_t)voff); dfree(ctx, d_cp, nsl); break; }
    if (bad) { dfree(ctx, vt, (size_t)voff); dfree(ctx, d_cp, nsl); continue; }      // hash collision: retry without sharing between slices
    dfree(ctx, m->col_ptr, nsl);
    m->col_ptr = d_cp; m->vt = vt; m->vt_len = voff; m->vshared_slices = shared; m->val_entries = entries;
    // scalar matrices: the code word most slices carry = the dominant stencil (kernel-parameter tables of the stencil kernel, spmv.cu)
    if (m->bb == 1 || m->bb == 9) {
      std::map<int64_t, int64_t> freq;
      for (size_t s = 0; s < nsl; s++) if (flag[s]) freq[ncp[s]]++;
      int64_t best = 0, bestn = 0;
      for (auto &kv : freq) if (kv.second > bestn) { best = kv.first; bestn = kv.second; }
      size_t s0 = 0;
      while (s0 < nsl && ncp[s0] != best) s0++;
      const int64_t w = s0 < nsl ? (sp[s0 + 1] - sp[s0]) >> 5 : 0;
      const char *mf = getenv("UGGPU_STENCIL_MIN_FRAC");          // tests: let a smaller share of the slices count as dominant
      const double minfrac = mf ? atof(mf) : 0.5;
      if (m->bb == 9) {
        if ((double)bestn > minfrac * (double)nsl && w == 27) {
          int32_t dist[27];
          Sten3 *s3 = new Sten3();
          memset(s3, 0, sizeof *s3);
          cudaMemcpyAsync(dist, m->col + UG_COLTAB(best), sizeof(int32_t) * 27, cudaMemcpyDeviceToHost, st);
          cudaMemcpyAsync(s3->v, vt + UG_VALTAB(best), sizeof(double) * 27 * 9, cudaMemcpyDeviceToHost, st);
          if (cudaStreamSynchronize(st) == cudaSuccess) {
            s3->code = best; s3->w = 27; s3->maxd = 0;
            for (int j = 0; j < 27; j++) { s3->dbytes[j] = (long long)dist[j] * 24; if (dist[j] > s3->maxd) s3->maxd = dist[j]; }
            m->sten3 = s3; m->sten_slices = bestn;
          } else delete s3;
        }
      } else if ((double)bestn > minfrac * (double)nsl && w >= 1 && w <= 32) {
        int32_t dist[32]; double vals[32];
        cudaMemcpyAsync(dist, m->col + UG_COLTAB(best), sizeof(int32_t) * w, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(vals, vt + UG_VALTAB(best), sizeof(double) * w, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) == cudaSuccess) {
          Sten sn; memset(&sn, 0, sizeof sn);
          sn.code = best; sn.w = (int)w; sn.maxd = 0;
          for (int j = 0; j < w; j++) { sn.dbytes[j] = (long long)dist[j] * 8; sn.v[j] = vals[j]; if (dist[j] > sn.maxd) sn.maxd = dist[j]; }
          m->sten = sn; m->sten_slices = bestn;
        }
      }
    }
    break;
  }
  dfree(ctx, d_hash, nsl + 1);
  return rc;
}

// ---- value dictionary of the transfer stencils ---------------------------------------------------------------------------------
#define VT_SLOTS 1024
#define VT_EMPTY 0xFFFFFFFFFFFFFFFFull
// open-addressing set of the bit patterns of all stored values; overflow[0] counts the distinct ones
__global__ void k_values_collect(int64_t count, const double *__restrict__ val, unsigned long long *tab, int *distinct)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val[i]);
    if (bits == VT_EMPTY) { atomicAdd(distinct, 100000); continue; }
    unsigned h = (unsigned)((bits * 0x9E3779B97F4A7C15ull) >> 54) & (VT_SLOTS - 1);
    for (int probe = 0; probe < VT_SLOTS; probe++) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tab + h);
      if (cur == bits) break;
      if (cur == VT_EMPTY) {
        cur = atomicCAS(tab + h, VT_EMPTY, bits);
        if (cur == VT_EMPTY) { atomicAdd(distinct, 1); break; }
        if (cur == bits) break;
      }
      h = (h + 1) & (VT_SLOTS - 1);
      if (probe == VT_SLOTS - 1) atomicAdd(distinct, 100000);
    }
    if (*reinterpret_cast<volatile int *>(distinct) > 256) return;       // too many: the caller gives up
  }
}

__global__ void k_values_encode(int64_t count, const double *__restrict__ val, const double *__restrict__ table, int nvals, uint8_t *__restrict__ code, int *bad)
{
  __shared__ unsigned long long st[256];
  for (int i = threadIdx.x; i < nvals; i += blockDim.x) st[i] = (unsigned long long)__double_as_longlong(table[i]);
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val[i]);
    int c = -1;
    for (int k = 0; k < nvals; k++) if (st[k] == bits) { c = k; break; }
    if (c < 0) { atomicAdd(bad, 1); c = 0; }
    code[i] = (uint8_t)c;
  }
}

int sell_compress_values(uggpu_ctx *ctx, SellMat *m)
{
  if (m->bb != 1 || m->padded <= 0 || m->vcode || getenv("UGGPU_NO_VALUE_TABLE")) return 0;
  cudaStream_t st = ctx->stream;
  unsigned long long *d_tab = nullptr; int *d_cnt = nullptr;
  UG_TRY(dalloc(ctx, &d_tab, (size_t)VT_SLOTS));
  UG_TRY(dalloc(ctx, &d_cnt, 2));
  CUDA_TRY(cudaMemsetAsync(d_tab, 0xff, sizeof(unsigned long long) * VT_SLOTS, st));
  CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), st));
  int blocks = (int)((m->padded + 255) / 256);
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  k_values_collect<<<blocks, 256, 0, st>>>(m->padded, m->val, d_tab, d_cnt);
  KCHECK(ctx);
  std::vector<unsigned long long> tab(VT_SLOTS);
  int cnt[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(tab.data(), d_tab, sizeof(unsigned long long) * VT_SLOTS, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  UG_TRY(dfree(ctx, d_tab, (size_t)VT_SLOTS));
  int rc = 0;
  if (cnt[0] >= 1 && cnt[0] <= 256) {
    std::vector<unsigned long long> vals;
    for (auto b : tab) if (b != VT_EMPTY) vals.push_back(b);
    std::sort(vals.begin(), vals.end());                 // deterministic codes
    if ((int)vals.size() == cnt[0]) {
      double table[256];
      for (int i = 0; i < 256; i++) { unsigned long long b = i < (int)vals.size() ? vals[i] : 0ull; memcpy(&table[i], &b, 8); }
      uint8_t *code = nullptr; double *d_table = nullptr;
      if ((rc = dalloc(ctx, &code, (size_t)m->padded)) == 0 && (rc = dalloc(ctx, &d_table, 256)) == 0) {
        cudaMemcpyAsync(d_table, table, sizeof table, cudaMemcpyHostToDevice, st);
        k_values_encode<<<blocks, 256, 0, st>>>(m->padded, m->val, d_table, (int)vals.size(), code, d_cnt + 1);
        ctx->launches++;
        int bad = 1;
        cudaMemcpyAsync(&bad, d_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "value table: %s", cudaGetErrorString(e));
        if (rc == 0 && bad == 0) { m->vcode = code; m->vtable = d_table; m->nvals = (int)vals.size(); }
        else { dfree(ctx, code, (size_t)m->padded); dfree(ctx, d_table, 256); }
      }
    }
  }
  dfree(ctx, d_cnt, 2);
  return rc;
}

__global__ void k_sell_diag(int n, int bb, const int64_t *__restrict__ slice_ptr, const uint16_t *__restrict__ rowlen, const double *__restrict__ val,
                            double *__restrict__ diag)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int nsl = (n + 31) >> 5;
  if ((r >> 5) >= nsl) return;
  const int s = r >> 5, lane = r & 31;
  const bool has = r < n && rowlen[r] > 0;
  const int64_t sp = slice_ptr[s];
  for (int k = 0; k < bb; k++) diag[((size_t)s * bb + k) * 32 + lane] = has ? val[(sp * bb) + (int64_t)k * 32 + lane] : 0.0;
}

int sell_update_diag(uggpu_ctx *ctx, SellMat *m)
{
  m->gen = ++ctx->value_gen;     // called after every change of the values

  if (m->n <= 0) return 0;
  const size_t nsl = (size_t)(m->n + 31) / 32;
  if (!m->diag) UG_TRY(dalloc(ctx, &m->diag, nsl * 32 * m->bb));
  k_sell_diag<<<(int)((nsl * 32 + 255) / 256), 256, 0, ctx->stream>>>(m->n, m->bb, m->slice_ptr, m->rowlen, m->val, m->diag);
  KCHECK(ctx);
  return 0;
}

int sell_free(uggpu_ctx *ctx, SellMat *m)
{
  if (m->n <= 0 && !m->col) { *m = SellMat(); return 0; }
  size_t nsl = (size_t)(m->n + 31) / 32;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  UG_TRY(sell_free_schedules(ctx, m));
  if (m->bnd_flag) dfree(ctx, m->bnd_flag, nsl);
  if (m->comm_flag) dfree(ctx, m->comm_flag, nsl + 1);
  stx_free(ctx, m);
  trc_free(ctx, m);
  if (m->vcode) dfree(ctx, m->vcode, (size_t)m->padded);
  if (m->vtable) dfree(ctx, m->vtable, 256);
  if (m->vt) dfree(ctx, m->vt, (size_t)m->vt_len);
  delete m->sten3;
  if (m->bnd_list) dfree(ctx, m->bnd_list, (size_t)(m->n_bnd > 0 ? m->n_bnd : 1));
  if (m->col_ptr != m->slice_ptr) dfree(ctx, m->col_ptr, nsl);
  dfree(ctx, m->slice_ptr, nsl + 1);
  dfree(ctx, m->rowlen, (size_t)m->n);
  dfree(ctx, m->col, (size_t)m->col_len);
  dfree(ctx, m->val, (size_t)m->padded * m->bb);
  if (m->diag) dfree(ctx, m->diag, nsl * 32 * m->bb);
  *m = SellMat();
  return 0;
}

// Deep copy of a matrix (pattern, layout, values, diagonal array) -- AllocMDFromMD + dmatcopy.  Schedules, interface lists and value
// tables are not copied (the first two are rebuilt on demand, the last exists for transfer stencils only).
int sell_clone(uggpu_ctx *ctx, const SellMat *src, SellMat *dst)
{
  if (src->vcode) return uggpu_fail(UGGPU_ERROR, "sell_clone: matrices with a value table cannot be copied");
  cudaStream_t st = ctx->stream;
  const size_t nsl = (size_t)(src->n + 31) / 32;
  SellMat m;
  m.n = src->n; m.bb = src->bb; m.nnz = src->nnz; m.padded = src->padded; m.maxlen = src->maxlen; m.fixed_w = src->fixed_w;
  m.col_len = src->col_len; m.uniform_slices = src->uniform_slices; m.col_words = src->col_words;
  int rc = 0;
#define CL(expr) do { if (!rc) rc = (expr); } while (0)
#define CC(expr) do { if (!rc) { cudaError_t e__ = (expr); if (e__ != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); } } while (0)
  CL(dalloc(ctx, &m.slice_ptr, nsl + 1));
  CL(dalloc(ctx, &m.rowlen, (size_t)m.n));
  CL(dalloc(ctx, &m.col, (size_t)m.col_len));
  CL(dalloc(ctx, &m.val, (size_t)m.padded * m.bb));
  if (!rc) {
    if (src->col_ptr != src->slice_ptr) CL(dalloc(ctx, &m.col_ptr, nsl)); else m.col_ptr = m.slice_ptr;
  }
  if (src->diag) CL(dalloc(ctx, &m.diag, nsl * 32 * m.bb));
  CC(cudaMemcpyAsync(m.slice_ptr, src->slice_ptr, sizeof(int64_t) * (nsl + 1), cudaMemcpyDeviceToDevice, st));
  CC(cudaMemcpyAsync(m.rowlen, src->rowlen, sizeof(uint16_t) * (size_t)m.n, cudaMemcpyDeviceToDevice, st));
  CC(cudaMemcpyAsync(m.col, src->col, sizeof(int32_t) * (size_t)m.col_len, cudaMemcpyDeviceToDevice, st));
  CC(cudaMemcpyAsync(m.val, src->val, sizeof(double) * (size_t)m.padded * m.bb, cudaMemcpyDeviceToDevice, st));
  if (!rc && src->col_ptr != src->slice_ptr) CC(cudaMemcpyAsync(m.col_ptr, src->col_ptr, sizeof(int64_t) * nsl, cudaMemcpyDeviceToDevice, st));
  if (!rc && src->diag) CC(cudaMemcpyAsync(m.diag, src->diag, sizeof(double) * nsl * 32 * m.bb, cudaMemcpyDeviceToDevice, st));
#undef CL
#undef CC
  if (rc) {
    if (m.col_ptr == m.slice_ptr) m.col_ptr = nullptr;
    if (m.col_ptr) dfree(ctx, m.col_ptr, nsl);
    if (m.slice_ptr) dfree(ctx, m.slice_ptr, nsl + 1);
    if (m.rowlen) dfree(ctx, m.rowlen, (size_t)m.n);
    if (m.col) dfree(ctx, m.col, (size_t)m.col_len);
    if (m.val) dfree(ctx, m.val, (size_t)m.padded * m.bb);
    if (m.diag) dfree(ctx, m.diag, nsl * 32 * m.bb);
    return rc;
  }
  *dst = m;
  return 0;
}

int sell_from_device_csr(uggpu_ctx *ctx, int n, int bb, const int64_t *d_rowptr, const int32_t *d_col, const double *d_val, SellMat *out)
{
  cudaStream_t st = ctx->stream;
  SellMat m;
  m.n = n; m.bb = bb;
  size_t nsl = (size_t)(n + 31) / 32;
  int *d_width = nullptr;
  UG_TRY(dalloc(ctx, &m.rowlen, (size_t)n));
  UG_TRY(dalloc(ctx, &m.slice_ptr, nsl + 1));
  UG_TRY(dalloc(ctx, &d_width, nsl));
  std::vector<int> width(nsl);
  std::vector<int64_t> sp;
  if (n > 0) {
    int blocks = (int)((nsl * 32 + 255) / 256);
    k_sell_rowlen<<<blocks, 256, 0, st>>>(n, d_rowptr, m.rowlen, d_width, ctx->derr);
    KCHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(width.data(), d_width, nsl * sizeof(int), cudaMemcpyDeviceToHost, st));
    int64_t last = 0;
    CUDA_TRY(cudaMemcpyAsync(&last, d_rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    UG_TRY(check_device_error(ctx));
    m.nnz = last;
  }
  sell_layout(width, sp, &m.maxlen, &m.fixed_w);
  m.padded = sp[nsl];
  UG_TRY(dfree(ctx, d_width, nsl));
  CUDA_TRY(cudaMemcpyAsync(m.slice_ptr, sp.data(), (nsl + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  UG_TRY(dalloc(ctx, &m.col, (size_t)m.padded));
  UG_TRY(dalloc(ctx, &m.val, (size_t)m.padded * bb));
  if (n > 0) {
    int blocks = (int)((nsl * 32 + 255) / 256);
    k_sell_fill<<<blocks, 256, 0, st>>>(n, bb, d_rowptr, d_col, d_val, m.slice_ptr, m.col, m.val);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaStreamSynchronize(st));   // sp (host vector) must outlive the copy
  m.col_ptr = m.slice_ptr; m.col_len = m.padded; m.col_words = m.nnz;
  UG_TRY(sell_compress_cols(ctx, &m));
  *out = m;
  return 0;
}

int sell_from_host_csr(uggpu_ctx *ctx, int n, int bb, const int32_t *rowptr, const int32_t *col, const double *val, SellMat *out)
{
  cudaStream_t st = ctx->stream;
  std::vector<int64_t> rp((size_t)n + 1);
  for (int i = 0; i <= n; i++) rp[i] = rowptr[i];
  int64_t nnz = n > 0 ? rp[n] : 0;
  if (nnz < 0) return uggpu_fail(UGGPU_ERROR, "negative nnz");
  int64_t *d_rp = nullptr; int32_t *d_col = nullptr; double *d_val = nullptr;
  UG_TRY(dalloc(ctx, &d_rp, (size_t)n + 1));
  UG_TRY(dalloc(ctx, &d_col, (size_t)nnz));
  UG_TRY(dalloc(ctx, &d_val, (size_t)nnz * bb));
  CUDA_TRY(cudaMemcpyAsync(d_rp, rp.data(), ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (nnz) {
    CUDA_TRY(cudaMemcpyAsync(d_col, col, (size_t)nnz * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if (val) CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nnz * bb * sizeof(double), cudaMemcpyHostToDevice, st));
    else CUDA_TRY(cudaMemsetAsync(d_val, 0, (size_t)nnz * bb * sizeof(double), st));        // pattern only
  }
  int rc = sell_from_device_csr(ctx, n, bb, d_rp, d_col, d_val, out);
  CUDA_TRY(cudaStreamSynchronize(st));
  dfree(ctx, d_rp, (size_t)n + 1); dfree(ctx, d_col, (size_t)nnz); dfree(ctx, d_val, (size_t)nnz * bb);
  return rc;
}

// rowlen -> int64 rowptr on the device (via host prefix sum; setup path only)
static int device_rowptr(uggpu_ctx *ctx, const SellMat *m, std::vector<int64_t> &rp, int64_t **d_rp)
{
  int n = m->n;
  std::vector<uint16_t> len((size_t)n);
  CUDA_TRY(cudaMemcpyAsync(len.data(), m->rowlen, (size_t)n * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  rp.assign((size_t)n + 1, 0);
  for (int i = 0; i < n; i++) rp[i + 1] = rp[i] + len[i];
  UG_TRY(dalloc(ctx, d_rp, (size_t)n + 1));
  CUDA_TRY(cudaMemcpyAsync(*d_rp, rp.data(), ((size_t)n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int sell_set_values_host(uggpu_ctx *ctx, SellMat *m, const double *val)
{
  if (m->vcode) { dfree(ctx, m->vcode, (size_t)m->padded); dfree(ctx, m->vtable, 256); m->nvals = 0; }     // stale codes
  UG_TRY(sell_free_schedules(ctx, m));     // they hold the old values; rebuilt on the next Gauss-Seidel solve
  std::vector<int64_t> rp; int64_t *d_rp = nullptr; double *d_val = nullptr;
  UG_TRY(device_rowptr(ctx, m, rp, &d_rp));
  size_t cnt = (size_t)m->nnz * m->bb;
  UG_TRY(dalloc(ctx, &d_val, cnt));
  CUDA_TRY(cudaMemcpyAsync(d_val, val, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (m->n > 0) {
    k_sell_set_values<<<(m->n + 255) / 256, 256, 0, ctx->stream>>>(m->n, m->bb, d_rp, d_val, m->slice_ptr, m->val);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_rp, (size_t)m->n + 1); dfree(ctx, d_val, cnt);
  return 0;
}

int sell_to_host_csr(uggpu_ctx *ctx, const SellMat *m, int32_t *rowptr, int32_t *col, double *val)
{
  std::vector<int64_t> rp; int64_t *d_rp = nullptr; int32_t *d_col = nullptr; double *d_val = nullptr;
  UG_TRY(device_rowptr(ctx, m, rp, &d_rp));
  if (rp[m->n] > 2147483647LL) return uggpu_fail(UGGPU_ERROR, "matrix has %lld entries: too many for the int32 CSR interface", (long long)rp[m->n]);
  if (rowptr) for (int i = 0; i <= m->n; i++) rowptr[i] = (int32_t)rp[i];
  size_t nnz = (size_t)m->nnz;
  if (col || val) {
    UG_TRY(dalloc(ctx, &d_col, nnz));
    UG_TRY(dalloc(ctx, &d_val, nnz * m->bb));
    if (m->n > 0) {
      k_sell_to_csr<<<(m->n + 255) / 256, 256, 0, ctx->stream>>>(view(*m), m->bb, d_rp, d_col, d_val);
      KCHECK(ctx);
    }
    if (col) CUDA_TRY(cudaMemcpyAsync(col, d_col, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (val) CUDA_TRY(cudaMemcpyAsync(val, d_val, nnz * m->bb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_rp, (size_t)m->n + 1); dfree(ctx, d_col, nnz); dfree(ctx, d_val, nnz * m->bb);
  return 0;
}
