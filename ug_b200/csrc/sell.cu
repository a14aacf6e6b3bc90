// sell.cu -- CSR/BSR (canonical, host-facing) <-> SELL-32 (device-resident) conversion.
//
// The canonical layout of SURVEY.md 8(a') -- rows in FIRSTVECTOR->SUCCVC order, entries in VSTART->MNEXT
// order -- is what crosses the C-ABI (uggpu_mat_set / uggpu_mat_get round-trip bit-exactly).  On the device the
// same entries are stored slice-interleaved so that thread-per-row kernels are coalesced (uggpu_internal.h).
#include "uggpu_internal.h"

#include <vector>

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

__global__ void k_sell_to_csr(int n, int bb, const int64_t *__restrict__ rowptr, const int64_t *__restrict__ slice_ptr,
                              const int32_t *__restrict__ col, const double *__restrict__ val, int32_t *__restrict__ ccol, double *__restrict__ cval)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int s = r >> 5, lane = r & 31;
  int64_t sp = slice_ptr[s];
  int64_t rp = rowptr[r];
  int len = (int)(rowptr[r + 1] - rp);
  for (int j = 0; j < len; j++) {
    int64_t src = sp + (int64_t)j * 32;
    ccol[rp + j] = col[src + lane];
    for (int k = 0; k < bb; k++) cval[(rp + j) * bb + k] = val[src * bb + (int64_t)k * 32 + lane];
  }
}

int sell_free(uggpu_ctx *ctx, SellMat *m)
{
  if (m->n <= 0 && !m->col) { *m = SellMat(); return 0; }
  size_t nsl = (size_t)(m->n + 31) / 32;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, m->slice_ptr, nsl + 1);
  dfree(ctx, m->rowlen, (size_t)m->n);
  dfree(ctx, m->col, (size_t)m->padded);
  dfree(ctx, m->val, (size_t)m->padded * m->bb);
  *m = SellMat();
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
  std::vector<int64_t> sp(nsl + 1, 0);
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
  for (size_t s = 0; s < nsl; s++) { sp[s + 1] = sp[s] + (int64_t)width[s] * 32; if (width[s] > m.maxlen) m.maxlen = width[s]; }
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
    CUDA_TRY(cudaMemcpyAsync(d_val, val, (size_t)nnz * bb * sizeof(double), cudaMemcpyHostToDevice, st));
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
      k_sell_to_csr<<<(m->n + 255) / 256, 256, 0, ctx->stream>>>(m->n, m->bb, d_rp, m->slice_ptr, m->col, m->val, d_col, d_val);
      KCHECK(ctx);
    }
    if (col) CUDA_TRY(cudaMemcpyAsync(col, d_col, nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (val) CUDA_TRY(cudaMemcpyAsync(val, d_val, nnz * m->bb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_rp, (size_t)m->n + 1); dfree(ctx, d_col, nnz); dfree(ctx, d_val, nnz * m->bb);
  return 0;
}
