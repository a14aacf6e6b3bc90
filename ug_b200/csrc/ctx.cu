// ctx.cu -- context, levels, flags, vectors of libuggpu.so (see include/uggpu.h).
#include "uggpu_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static thread_local std::string g_last_error;

int uggpu_fail(int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

extern "C" const char *uggpu_last_error(void) { return g_last_error.c_str(); }

int dev_alloc(uggpu_ctx *ctx, void **p, size_t bytes)
{
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return uggpu_fail(UGGPU_OUT_OF_MEM, "cudaMalloc(%zu bytes) failed: %s (context holds %lld bytes)", bytes, cudaGetErrorString(e), (long long)ctx->bytes);
  }
  ctx->bytes += (int64_t)bytes;
  return 0;
}

int dev_free(uggpu_ctx *ctx, void *p, size_t bytes)
{
  if (p == nullptr) return 0;
  if (bytes == 0) bytes = 16;
  CUDA_TRY(cudaFree(p));
  ctx->bytes -= (int64_t)bytes;
  return 0;
}

Level *get_level(uggpu_ctx *ctx, int level)
{
  if (ctx == nullptr) { uggpu_fail(UGGPU_ERROR, "null context"); return nullptr; }
  if (level < 0 || level >= UGGPU_MAX_LEVELS || !ctx->lev[level].exists) { uggpu_fail(UGGPU_ERROR, "level %d does not exist", level); return nullptr; }
  return &ctx->lev[level];
}

double *get_vec_lazy(uggpu_ctx *ctx, int level, int vec)
{
  Level *L = get_level(ctx, level);
  if (!L) return nullptr;
  auto it = L->vecs.find(vec);
  if (it == L->vecs.end()) { uggpu_fail(UGGPU_DESC_MISMATCH, "vector %d not allocated on level %d", vec, level); return nullptr; }
  return it->second;
}

// compute stream waits for an upload of the vector that is still in flight on the copy stream (uggpu_vec_upload_async)
int vec_wait(uggpu_ctx *ctx, int level, int vec)
{
  Level *L = &ctx->lev[level];
  auto it = L->pending.find(vec);
  if (it == L->pending.end()) return 0;
  CUDA_TRY(cudaStreamWaitEvent(ctx->stream, it->second, 0));
  CUDA_TRY(cudaEventDestroy(it->second));
  L->pending.erase(it);
  return 0;
}

double *get_vec(uggpu_ctx *ctx, int level, int vec)
{
  double *p = get_vec_lazy(ctx, level, vec);
  if (p && !ctx->lev[level].pending.empty() && vec_wait(ctx, level, vec)) return nullptr;
  return p;
}

SellMat *get_mat(uggpu_ctx *ctx, int level, int mat)
{
  Level *L = get_level(ctx, level);
  if (!L) return nullptr;
  auto it = L->mats.find(mat);
  if (it == L->mats.end()) { uggpu_fail(UGGPU_DESC_MISMATCH, "matrix %d not set on level %d", mat, level); return nullptr; }
  return &it->second;
}

SellMat *get_mat_quiet(uggpu_ctx *ctx, int level, int mat)      // nullptr without an error message when the matrix does not exist
{
  if (level < 0 || level >= UGGPU_MAX_LEVELS || !ctx->lev[level].exists) return nullptr;
  auto it = ctx->lev[level].mats.find(mat);
  return it == ctx->lev[level].mats.end() ? nullptr : &it->second;
}

// Distance rule (measured on B200, 513^3, profiles/README.md): 3/4 of the warps resident on the whole GPU, i.e. the slice a
// warp of the next generation will start with -- but never more than ~40 MB of matrix ahead (beyond ~60 MB the prefetched
// lines are evicted from the 126 MB L2 before they are used and the kernel gets slower than without prefetch), and off when
// that cap falls below half a generation (3x3 blocks with 27 entries: 62 KB per slice; such rows are long streams per thread
// and reach 0.9 of the HBM peak without help).
Prefetch make_prefetch(const uggpu_ctx *ctx, const SellMat *A, int bs, int slices_per_warp)
{
  Prefetch pf;
  const char *d = getenv("UGGPU_PF_DIST"), *m = getenv("UGGPU_PF_MODE");
  const int64_t resident = (int64_t)ctx->sm_count * (2048 / 32);
  // slices with shared value tables (sell_share_values) have no value stream: what a slice pulls through L2 is its vector rows
  const bool vshared = A->vt && A->vshared_slices * 2 > (int64_t)(A->n + 31) / 32;
  const int64_t slice_bytes = vshared ? (int64_t)1024 * bs : (int64_t)(A->maxlen > 0 ? A->maxlen : 1) * A->bb * 256;
  int64_t dist = resident * 3 / 4 * slices_per_warp;      // a resident warp holds slices_per_warp slices: one generation is that much longer
  const int64_t cap = ((int64_t)40 << 20) / slice_bytes;
  if (cap < dist) dist = cap;
  if (dist < resident / 2 * slices_per_warp) dist = 0;
  pf.dist = d ? atoi(d) : (int)dist;
  pf.mode = m ? atoi(m) : 63;
  pf.nsl = (A->n + 31) / 32;
  pf.val_lines = (A->maxlen * A->bb * 256 + 127) / 128;
  pf.col_lines = A->maxlen < 32 ? A->maxlen : 32;
  pf.val_bytes = A->padded * A->bb * (int64_t)sizeof(double);
  pf.col_bytes = A->col_len * (int64_t)sizeof(int32_t);
  pf.vec_bytes = (int64_t)A->n * bs * (int64_t)sizeof(double);     // vectors indexed by the matrix' rows
  return pf;
}

int ensure_partials(uggpu_ctx *ctx, size_t count)
{
  if (count <= ctx->partials_cap) return 0;
  if (ctx->partials) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); UG_TRY(dfree(ctx, ctx->partials, ctx->partials_cap)); }
  size_t cap = count + count / 2 + 1024;
  UG_TRY(dalloc(ctx, &ctx->partials, cap));
  ctx->partials_cap = cap;
  return 0;
}

int check_device_error(uggpu_ctx *ctx)
{
  CUDA_TRY(cudaMemcpyAsync(ctx->herr, ctx->derr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int e = *ctx->herr;
  if (e) {
    CUDA_TRY(cudaMemsetAsync(ctx->derr, 0, sizeof(int), ctx->stream));
    return uggpu_fail(e, "device kernel reported error %d (%s)", e, e == UGGPU_SMALL_DIAG ? "NUM_SMALL_DIAG: singular diagonal block" : "see code");
  }
  return 0;
}

#define RES_SLOTS (UGGPU_MAX_LEVELS * 4 * UGGPU_MAX_BS)

extern "C" int uggpu_ctx_create(int device, uggpu_ctx **out)
{
  if (out == nullptr) return uggpu_fail(UGGPU_ERROR, "null out pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return uggpu_fail(UGGPU_CUDA_ERROR, "no usable CUDA device (%s); libuggpu has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  }
  if (device < 0 || device >= ndev) return uggpu_fail(UGGPU_ERROR, "device %d out of range (0..%d)", device, ndev - 1);
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return uggpu_fail(UGGPU_CUDA_ERROR, "device %d is sm_%d%d; libuggpu is built for sm_100a only", device, prop.major, prop.minor);
  uggpu_ctx *ctx = new uggpu_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  UG_TRY(dalloc(ctx, &ctx->dres, RES_SLOTS));
  UG_TRY(dalloc(ctx, &ctx->derr, 4));
  CUDA_TRY(cudaMemsetAsync(ctx->derr, 0, 4 * sizeof(int), ctx->stream));
  CUDA_TRY(cudaMallocHost(&ctx->hres, RES_SLOTS * sizeof(double)));
  CUDA_TRY(cudaMallocHost(&ctx->herr, 4 * sizeof(int)));
  *out = ctx;
  return 0;
}

extern "C" int uggpu_ctx_destroy(uggpu_ctx *ctx)
{
  if (ctx == nullptr) return 0;
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); ctx->copy_stream = nullptr; }
  if (ctx->halo_stream) { cudaStreamSynchronize(ctx->halo_stream); cudaStreamDestroy(ctx->halo_stream); ctx->halo_stream = nullptr; }
  for (int i = 0; i < 2; i++) if (ctx->halo_ev[i]) { cudaEventDestroy(ctx->halo_ev[i]); ctx->halo_ev[i] = nullptr; }
  uggpu_comm_destroy(ctx);
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++)
    if (ctx->lev[l].exists) uggpu_level_destroy(ctx, l);
  if (ctx->partials) dfree(ctx, ctx->partials, ctx->partials_cap);
  dfree(ctx, ctx->dres, RES_SLOTS);
  dfree(ctx, ctx->derr, 4);
  cudaFreeHost(ctx->hres);
  cudaFreeHost(ctx->herr);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

extern "C" int uggpu_sync(uggpu_ctx *ctx)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int uggpu_stream(uggpu_ctx *ctx, void **stream)
{
  if (!ctx || !stream) return uggpu_fail(UGGPU_ERROR, "null argument");
  *stream = (void *)ctx->stream;
  return 0;
}

extern "C" int uggpu_set_fullrefinelevel(uggpu_ctx *ctx, int level)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  ctx->fullrefinelevel = level;
  return 0;
}

extern "C" int64_t uggpu_launch_count(uggpu_ctx *ctx) { return ctx ? ctx->launches : -1; }
extern "C" int64_t uggpu_device_bytes(uggpu_ctx *ctx) { return ctx ? ctx->bytes : -1; }

// ---- levels ---------------------------------------------------------------------------------------------
__global__ void k_fill_u8(uint8_t *p, size_t n, uint8_t v)
{
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

static int fill_u8(uggpu_ctx *ctx, uint8_t *p, size_t n, uint8_t v)
{
  CUDA_TRY(cudaMemsetAsync(p, v, n, ctx->stream));
  return 0;
}

extern "C" int uggpu_level_create(uggpu_ctx *ctx, int level, int n, int bs)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  if (level < 0 || level >= UGGPU_MAX_LEVELS) return uggpu_fail(UGGPU_ERROR, "level %d out of range", level);
  if (bs < 1 || bs > UGGPU_MAX_BS) return uggpu_fail(UGGPU_BLOCK_TOO_LARGE, "block size %d not in 1..%d", bs, UGGPU_MAX_BS);
  if (n < 0) return uggpu_fail(UGGPU_ERROR, "negative n");
  CUDA_TRY(cudaSetDevice(ctx->device));
  if (ctx->lev[level].exists) UG_TRY(uggpu_level_destroy(ctx, level));
  Level &L = ctx->lev[level];
  L.n = n; L.bs = bs;
  UG_TRY(dalloc(ctx, &L.vclass, (size_t)n));
  UG_TRY(dalloc(ctx, &L.vnclass, (size_t)n));
  UG_TRY(dalloc(ctx, &L.ctl, (size_t)n));
  UG_TRY(dalloc(ctx, &L.skip, (size_t)n));
  L.exists = true;
  return uggpu_level_set_flags(ctx, level, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int uggpu_level_destroy(uggpu_ctx *ctx, int level)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  size_t n = (size_t)L->n;
  for (auto &kv : L->mats) sell_free(ctx, &kv.second);
  for (auto &kv : L->pending) { cudaEventSynchronize(kv.second); cudaEventDestroy(kv.second); }
  L->pending.clear();
  for (auto &kv : L->vecs) { double *p = kv.second; if (!halo_vec_release(ctx, L, p, (vec_count(L) ? vec_count(L) : 2) * sizeof(double))) dfree(ctx, p, vec_count(L)); }
  level_free_part(ctx, L);
  sell_free(ctx, &L->P);
  sell_free(ctx, &L->R);
  level_free_lu(ctx, L);
  dfree(ctx, L->vclass, n); dfree(ctx, L->vnclass, n); dfree(ctx, L->ctl, n); dfree(ctx, L->skip, n);
  *L = Level();
  return 0;
}

extern "C" int uggpu_level_n(uggpu_ctx *ctx, int level) { Level *L = get_level(ctx, level); return L ? L->n : -1; }
extern "C" int uggpu_level_bs(uggpu_ctx *ctx, int level) { Level *L = get_level(ctx, level); return L ? L->bs : -1; }

extern "C" int uggpu_level_set_flags(uggpu_ctx *ctx, int level, const uint8_t *vclass, const uint8_t *vnclass, const uint8_t *ctl, const uint32_t *skip)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  size_t n = (size_t)L->n;
  if (n == 0) return 0;
  cudaStream_t s = ctx->stream;
  if (vclass) CUDA_TRY(cudaMemcpyAsync(L->vclass, vclass, n, cudaMemcpyHostToDevice, s)); else UG_TRY(fill_u8(ctx, L->vclass, n, 3));
  if (vnclass) CUDA_TRY(cudaMemcpyAsync(L->vnclass, vnclass, n, cudaMemcpyHostToDevice, s)); else UG_TRY(fill_u8(ctx, L->vnclass, n, 3));
  if (ctl) CUDA_TRY(cudaMemcpyAsync(L->ctl, ctl, n, cudaMemcpyHostToDevice, s)); else UG_TRY(fill_u8(ctx, L->ctl, n, UGGPU_CTL_NEW_DEFECT | UGGPU_CTL_FINE_GRID_DOF));
  if (skip) CUDA_TRY(cudaMemcpyAsync(L->skip, skip, n * 4, cudaMemcpyHostToDevice, s)); else CUDA_TRY(cudaMemsetAsync(L->skip, 0, n * 4, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // host buffers may be reused by the caller
  return 0;
}

extern "C" int uggpu_level_get_flags(uggpu_ctx *ctx, int level, uint8_t *vclass, uint8_t *vnclass, uint8_t *ctl, uint32_t *skip)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  size_t n = (size_t)L->n;
  cudaStream_t s = ctx->stream;
  if (vclass) CUDA_TRY(cudaMemcpyAsync(vclass, L->vclass, n, cudaMemcpyDeviceToHost, s));
  if (vnclass) CUDA_TRY(cudaMemcpyAsync(vnclass, L->vnclass, n, cudaMemcpyDeviceToHost, s));
  if (ctl) CUDA_TRY(cudaMemcpyAsync(ctl, L->ctl, n, cudaMemcpyDeviceToHost, s));
  if (skip) CUDA_TRY(cudaMemcpyAsync(skip, L->skip, n * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return 0;
}

// ---- matrices ------------------------------------------------------------------------------------------------
extern "C" int uggpu_mat_set(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col, const double *val)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!rowptr || !col || !val) return uggpu_fail(UGGPU_ERROR, "uggpu_mat_set: null array");
  auto it = L->mats.find(mat);
  if (it != L->mats.end()) { UG_TRY(sell_free(ctx, &it->second)); L->mats.erase(it); }
  SellMat m;
  UG_TRY(sell_from_host_csr(ctx, L->n, L->bs * L->bs, rowptr, col, val, &m));
  UG_TRY(sell_update_diag(ctx, &m));
  UG_TRY(sell_share_values(ctx, &m));
  L->mats[mat] = m;
  return 0;
}

extern "C" int uggpu_mat_set_pattern(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!rowptr || !col) return uggpu_fail(UGGPU_ERROR, "uggpu_mat_set_pattern: null array");
  auto it = L->mats.find(mat);
  if (it != L->mats.end()) { UG_TRY(sell_free(ctx, &it->second)); L->mats.erase(it); }
  SellMat m;
  UG_TRY(sell_from_host_csr(ctx, L->n, L->bs * L->bs, rowptr, col, nullptr, &m));     // val == nullptr: zeros
  UG_TRY(sell_update_diag(ctx, &m));
  L->mats[mat] = m;
  return 0;
}

extern "C" int uggpu_mat_set_values(uggpu_ctx *ctx, int level, int mat, const double *val)
{
  SellMat *m = get_mat(ctx, level, mat);
  if (!m) return UGGPU_DESC_MISMATCH;
  UG_TRY(sell_set_values_host(ctx, m, val));
  UG_TRY(sell_update_diag(ctx, m));
  return sell_share_values(ctx, m);        // the tables follow the new values
}

extern "C" int uggpu_mat_get(uggpu_ctx *ctx, int level, int mat, int32_t *rowptr, int32_t *col, double *val)
{
  SellMat *m = get_mat(ctx, level, mat);
  if (!m) return UGGPU_DESC_MISMATCH;
  return sell_to_host_csr(ctx, m, rowptr, col, val);
}

extern "C" int64_t uggpu_mat_nnz(uggpu_ctx *ctx, int level, int mat)
{
  SellMat *m = get_mat(ctx, level, mat);
  return m ? m->nnz : -1;
}

extern "C" int64_t uggpu_mat_col_words(uggpu_ctx *ctx, int level, int mat)
{
  SellMat *m = get_mat(ctx, level, mat);
  return m ? m->col_words : -1;
}

extern "C" int64_t uggpu_mat_val_entries(uggpu_ctx *ctx, int level, int mat)
{
  SellMat *m = get_mat(ctx, level, mat);
  return m ? (m->val_entries >= 0 ? m->val_entries : m->nnz) : -1;
}

extern "C" int64_t uggpu_mat_stencil_slices(uggpu_ctx *ctx, int level, int mat)
{
  SellMat *m = get_mat(ctx, level, mat);
  return m ? ((m->sten.w > 0 || m->sten3) ? m->sten_slices : 0) : -1;
}

extern "C" double uggpu_mat_pass_bytes(uggpu_ctx *ctx, int level, int mat)
{
  Level *L = get_level(ctx, level);
  SellMat *m = get_mat(ctx, level, mat);
  if (!L || !m) return -1.0;
  const double sb = stx_matrix_bytes(L, m);
  return sb >= 0 ? sb : m->entry_bytes() + 4.0 * (L->n + 1.0);
}

extern "C" int64_t uggpu_mat_padded_nnz(uggpu_ctx *ctx, int level, int mat)
{
  SellMat *m = get_mat(ctx, level, mat);
  return m ? m->padded : -1;
}

extern "C" int uggpu_mat_free(uggpu_ctx *ctx, int level, int mat)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  auto it = L->mats.find(mat);
  if (it == L->mats.end()) return 0;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  UG_TRY(sell_free(ctx, &it->second));
  L->mats.erase(it);
  return 0;
}

extern "C" int uggpu_transfer_set(uggpu_ctx *ctx, int level, const int32_t *p_rowptr, const int32_t *p_col, const double *p_w,
                                  const int32_t *r_rowptr, const int32_t *r_col, const double *r_w)
{
  Level *L = get_level(ctx, level);
  Level *C = get_level(ctx, level - 1);
  if (!L || !C) return UGGPU_NO_COARSER_GRID;
  if (!p_rowptr || !p_col || !p_w || !r_rowptr || !r_col || !r_w) return uggpu_fail(UGGPU_ERROR, "uggpu_transfer_set: null array");
  UG_TRY(sell_free(ctx, &L->P));
  UG_TRY(sell_free(ctx, &L->R));
  UG_TRY(sell_from_host_csr(ctx, L->n, 1, p_rowptr, p_col, p_w, &L->P));
  UG_TRY(sell_from_host_csr(ctx, C->n, 1, r_rowptr, r_col, r_w, &L->R));
  UG_TRY(sell_compress_values(ctx, &L->P));
  UG_TRY(sell_compress_values(ctx, &L->R));
  return 0;
}

extern "C" int uggpu_transfer_get(uggpu_ctx *ctx, int level, int which, int32_t *rowptr, int32_t *col, double *w)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  SellMat *m = which == 0 ? &L->P : &L->R;
  if (!m->valid()) return uggpu_fail(UGGPU_NO_COARSER_GRID, "no transfer set on level %d", level);
  return sell_to_host_csr(ctx, m, rowptr, col, w);
}

extern "C" int64_t uggpu_transfer_nnz(uggpu_ctx *ctx, int level, int which)
{
  Level *L = get_level(ctx, level);
  if (!L) return -1;
  SellMat *m = which == 0 ? &L->P : &L->R;
  return m->valid() ? m->nnz : -1;
}

// ---- vectors ------------------------------------------------------------------------------------------------------
extern "C" int uggpu_vec_alloc(uggpu_ctx *ctx, int level, int vec)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (L->vecs.count(vec)) return 0;
  double *p = nullptr;
  size_t cnt = vec_count(L);
  UG_TRY(dalloc(ctx, &p, cnt));
  CUDA_TRY(cudaMemsetAsync(p, 0, (cnt ? cnt : 2) * sizeof(double), ctx->stream));
  L->vecs[vec] = p;
  return 0;
}

extern "C" int uggpu_vec_free(uggpu_ctx *ctx, int level, int vec)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  auto it = L->vecs.find(vec);
  if (it == L->vecs.end()) return 0;
  UG_TRY(vec_wait(ctx, level, vec));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  double *p = it->second;
  if (!halo_vec_release(ctx, L, p, (vec_count(L) ? vec_count(L) : 2) * sizeof(double))) UG_TRY(dfree(ctx, p, vec_count(L)));
  L->vecs.erase(it);
  return 0;
}

extern "C" int uggpu_vec_upload(uggpu_ctx *ctx, int level, int vec, const double *host)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!L->vecs.count(vec)) UG_TRY(uggpu_vec_alloc(ctx, level, vec));
  UG_TRY(vec_wait(ctx, level, vec));
  double *p = L->vecs[vec];
  CUDA_TRY(cudaMemcpyAsync(p, host, (size_t)L->n * L->bs * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Upload that overlaps whatever the compute stream is doing: the copy runs on a second stream behind the work already
// enqueued (so it cannot overtake kernels that still use the vector) and the first later operation that touches the
// vector waits for it.  `host` must stay valid until then (pinned memory for a truly asynchronous copy).  Used for the
// iterate x of a solve: only the last kernel of the first cycle needs it.
extern "C" int uggpu_vec_upload_async(uggpu_ctx *ctx, int level, int vec, const double *host)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!L->vecs.count(vec)) UG_TRY(uggpu_vec_alloc(ctx, level, vec));
  UG_TRY(vec_wait(ctx, level, vec));
  if (!ctx->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  cudaEvent_t before, done;
  CUDA_TRY(cudaEventCreateWithFlags(&before, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventRecord(before, ctx->stream));
  CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, before, 0));
  CUDA_TRY(cudaEventDestroy(before));
  CUDA_TRY(cudaMemcpyAsync(L->vecs[vec], host, (size_t)L->n * L->bs * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
  CUDA_TRY(cudaEventRecord(done, ctx->copy_stream));
  L->pending[vec] = done;
  return 0;
}

extern "C" int uggpu_vec_download(uggpu_ctx *ctx, int level, int vec, double *host)
{
  double *p = get_vec(ctx, level, vec);
  if (!p) return UGGPU_DESC_MISMATCH;
  Level *L = &ctx->lev[level];
  CUDA_TRY(cudaMemcpyAsync(host, p, (size_t)L->n * L->bs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int uggpu_vec_devptr(uggpu_ctx *ctx, int level, int vec, void **dptr)
{
  double *p = get_vec(ctx, level, vec);
  if (!p) return UGGPU_DESC_MISMATCH;
  *dptr = p;
  return 0;
}

// ---- per-kernel profiling ------------------------------------------------------------------------------------------
static cudaEvent_t prof_event(uggpu_ctx *ctx)
{
  cudaEvent_t e;
  if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(uggpu_ctx *c, int kind, int level, double bytes) : ctx(c), on(c->prof)
{
  if (!on) return;
  rec.kind = kind; rec.level = level; rec.bytes = bytes;
  rec.e0 = prof_event(ctx); rec.e1 = prof_event(ctx);
  cudaEventRecord(rec.e0, ctx->stream);
}

ProfScope::~ProfScope()
{
  if (!on) return;
  cudaEventRecord(rec.e1, ctx->stream);
  ctx->prof_recs.push_back(rec);
}

extern "C" int uggpu_prof_enable(uggpu_ctx *ctx, int on)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (auto &r : ctx->prof_recs) { ctx->prof_pool.push_back(r.e0); ctx->prof_pool.push_back(r.e1); }
  ctx->prof_recs.clear();
  ctx->prof = on != 0;
  return 0;
}

extern "C" int uggpu_prof_summary(uggpu_ctx *ctx, int kind, int level, int64_t *launches, double *ms, double *alg_bytes)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int64_t cnt = 0; double t = 0.0, by = 0.0;
  for (auto &r : ctx->prof_recs) {
    if ((kind >= 0 && r.kind != kind) || (level >= 0 && r.level != level)) continue;
    float e = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&e, r.e0, r.e1));
    cnt++; t += e; by += r.bytes;
  }
  if (launches) *launches = cnt;
  if (ms) *ms = t;
  if (alg_bytes) *alg_bytes = by;
  return 0;
}
