// dataio.cu -- savedata / loaddata for device vectors: np/udm/data_io.cc:650 SaveData, :408 LoadData.
//
// The reference writes one record of ncomp doubles per NODE in the order of the node IDs (all levels), the components of the saved
// VECDATA_DESCs side by side, behind a header (np/udm/dio.cc:338 Write_DT_General) whose first two items are always ASCII and whose
// remaining items and the body follow the file's mode (low/bio.cc: "asc" -- `%d\n`, `%g\n`, strings as `len\n` + characters + blank;
// "bin" -- raw ints and doubles, strings as `len ` + characters + blank).  Here the body is gathered ON THE DEVICE from the vectors of
// all levels into one buffer in node-ID order (k_data_pack: one thread per value), copied down once and written by the same format
// rules -- files are byte-identical to the reference's (the ASCII mode keeps its 6 significant digits, like the reference).  Loading is
// the reverse: parse, upload once, scatter (k_data_unpack).  uggpu_data_write / uggpu_data_read are the host-only halves (no device
// needed): they are what the CPU tests pin against the reference's files.
#include "uggpu_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define DIO_TITLE "####.sparse.data.storage.format.####"      // np/udm/dio.cc:61
#define DIO_VERSION_STR "DATA_IO_1.7"                         // np/udm/dio.h:48
enum { MODE_ASCII = 1, MODE_BIN = 2 };                        // low/bio.h BIO_ASCII / BIO_BIN

namespace {
struct Writer {
  FILE *f; int mode;
  bool str(const char *s) { return fprintf(f, mode == MODE_ASCII ? "%d\n" : "%d ", (int)strlen(s)) >= 0 && fputs(s, f) >= 0 && fputc(' ', f) != EOF; }
  bool ints(int n, const int *v)
  {
    if (mode == MODE_BIN) return fwrite(v, sizeof(int) * n, 1, f) == 1;
    for (int i = 0; i < n; i++) if (fprintf(f, "%d\n", v[i]) < 0) return false;
    return true;
  }
  bool doubles(size_t n, const double *v)
  {
    if (n == 0) return true;
    if (mode == MODE_BIN) return fwrite(v, sizeof(double) * n, 1, f) == 1;
    for (size_t i = 0; i < n; i++) if (fprintf(f, "%g\n", v[i]) < 0) return false;
    return true;
  }
};
struct Reader {
  FILE *f; int mode;
  bool str(std::string &out)
  {
    int len;
    if (fscanf(f, mode == MODE_ASCII ? "%d\n" : "%d ", &len) != 1 || len < 0 || len > 8192) return false;
    out.resize((size_t)len);
    for (int i = 0; i < len; i++) { int c = fgetc(f); if (c == EOF) return false; out[i] = (char)c; }
    return fgetc(f) == ' ';
  }
  bool ints(int n, int *v)
  {
    if (mode == MODE_BIN) return fread(v, sizeof(int) * n, 1, f) == 1;
    for (int i = 0; i < n; i++) if (fscanf(f, "%d\n", v + i) != 1) return false;
    return true;
  }
  bool doubles(size_t n, double *v)
  {
    if (n == 0) return true;
    if (mode == MODE_BIN) return fread(v, sizeof(double) * n, 1, f) == 1;
    for (size_t i = 0; i < n; i++) if (fscanf(f, "%lg\n", v + i) != 1) return false;
    return true;
  }
};
int mode_of(const char *type)
{
  if (type && strcmp(type, "asc") == 0) return MODE_ASCII;
  if (type && strcmp(type, "bin") == 0) return MODE_BIN;
  return 0;                                                   // "xdr" (data_io.cc:716) is not offered
}
}  // namespace

extern "C" int uggpu_data_write(const char *filename, const char *type, const uggpu_data_general *g, int nvd, const int *ncomp,
                                const char *const *vdname, const char *const *compnames, int64_t nnode, const double *data)
{
  const int mode = mode_of(type);
  if (!mode) return uggpu_fail(UGGPU_ERROR, "uggpu_data_write: type '%s' (asc | bin)", type ? type : "");
  if (!filename || !g || nvd < 1 || nvd > 100 || !ncomp || !vdname || !compnames) return uggpu_fail(UGGPU_ERROR, "uggpu_data_write: bad argument");
  int total = 0;
  for (int i = 0; i < nvd; i++) total += ncomp[i];
  if ((int64_t)total * nnode > 2147483647LL) return uggpu_fail(UGGPU_ERROR, "uggpu_data_write: ndata does not fit the format's int");
  FILE *f = fopen(filename, "w");
  if (!f) return uggpu_fail(UGGPU_ERROR, "uggpu_data_write: cannot open %s", filename);
  Writer w{f, MODE_ASCII};
  bool ok = w.str(DIO_TITLE) && w.ints(1, &mode);              // head always in ASCII (dio.cc:345)
  w.mode = mode;
  ok = ok && w.str(DIO_VERSION_STR) && w.str(g->ident ? g->ident : "---") && w.str(g->mgfile ? g->mgfile : "saved_without_mg");
  ok = ok && w.doubles(1, &g->time) && w.doubles(1, &g->dt) && w.doubles(1, &g->ndt);
  const int four[4] = {g->nparfiles, g->me, g->magic_cookie, nvd};
  ok = ok && w.ints(4, four);
  for (int i = 0; ok && i < nvd; i++) {
    const int vdtype = ncomp[i] == 1 ? 0 : 2;                  // DIO_SCALAR / DIO_MULTIPLE_SCALAR (data_io.cc:779)
    ok = w.str(vdname[i]) && w.ints(1, ncomp + i) && w.ints(1, &vdtype) && w.str(compnames[i]);
  }
  const int ndata = (int)(total * nnode);
  ok = ok && w.ints(1, &ndata) && w.doubles((size_t)ndata, data);
  if (fclose(f)) ok = false;
  return ok ? 0 : uggpu_fail(UGGPU_ERROR, "uggpu_data_write: write error on %s", filename);
}

// header of a data file (Read_DT_General dio.cc:273); *body = file offset of the first value
static int read_header(FILE *f, Reader &r, uggpu_data_general *g, std::string &ident, std::string &mgfile, int *nvd, std::vector<int> &ncomp,
                       std::vector<std::string> &names, int *ndata)
{
  std::string s;
  r.mode = MODE_ASCII;
  int mode = 0;
  if (!r.str(s) || s != DIO_TITLE || !r.ints(1, &mode) || (mode != MODE_ASCII && mode != MODE_BIN)) return 1;
  r.mode = mode;
  if (!r.str(s) || s != DIO_VERSION_STR) return 2;           // "wrong version" (data_io.cc:508); 1.6 files (no ident) are not accepted
  int four[4];
  if (!r.str(ident) || !r.str(mgfile) || !r.doubles(1, &g->time) || !r.doubles(1, &g->dt) || !r.doubles(1, &g->ndt) || !r.ints(4, four)) return 1;
  g->nparfiles = four[0]; g->me = four[1]; g->magic_cookie = four[2];
  *nvd = four[3];
  if (*nvd < 0 || *nvd > 100) return 1;
  ncomp.assign((size_t)*nvd, 0); names.assign((size_t)*nvd, std::string());
  for (int i = 0; i < *nvd; i++) {
    int vdtype;
    if (!r.str(names[i]) || !r.ints(1, &ncomp[i]) || !r.ints(1, &vdtype) || !r.str(s)) return 1;
  }
  return r.ints(1, ndata) ? 0 : 1;
}

extern "C" int uggpu_data_read(const char *filename, uggpu_data_general *g, int *nvd, int *ncomp, int ncomp_cap, int64_t *ndata, double *data, int64_t data_cap)
{
  if (!filename || !g || !nvd || !ndata) return uggpu_fail(UGGPU_ERROR, "uggpu_data_read: null argument");
  FILE *f = fopen(filename, "r");
  if (!f) return uggpu_fail(UGGPU_ERROR, "uggpu_data_read: cannot open %s", filename);
  Reader r{f, MODE_ASCII};
  static thread_local std::string ident, mgfile;
  std::vector<int> nc; std::vector<std::string> names;
  int nd = 0;
  int rc = read_header(f, r, g, ident, mgfile, nvd, nc, names, &nd);
  if (rc) { fclose(f); return uggpu_fail(UGGPU_ERROR, rc == 2 ? "uggpu_data_read: %s: wrong version" : "uggpu_data_read: %s is not a data file", filename); }
  g->ident = ident.c_str(); g->mgfile = mgfile.c_str();       // valid until the next call on this thread
  *ndata = nd;
  if (ncomp) for (int i = 0; i < *nvd && i < ncomp_cap; i++) ncomp[i] = nc[i];
  bool ok = true;
  if (data) {
    if (data_cap < nd) { fclose(f); return uggpu_fail(UGGPU_ERROR, "uggpu_data_read: buffer of %lld values for %d", (long long)data_cap, nd); }
    ok = r.doubles((size_t)nd, data);
  }
  fclose(f);
  return ok ? 0 : uggpu_fail(UGGPU_ERROR, "uggpu_data_read: %s is truncated", filename);
}

// ---- device side: body <-> vectors of all levels --------------------------------------------------------------------------------------
// vp[v * UGGPU_MAX_LEVELS + level] = vector v on that level (nullptr: descriptor skipped on load)
__global__ void k_data_pack(int64_t nnode, int nvd, int bs, const int32_t *__restrict__ id_level, const int32_t *__restrict__ id_row,
                            const double *const *__restrict__ vp, double *__restrict__ body)
{
  const int total = nvd * bs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnode * total) return;
  const int64_t id = i / total;
  const int s = (int)(i - id * total), v = s / bs, c = s - v * bs;
  body[i] = vp[v * UGGPU_MAX_LEVELS + id_level[id]][(size_t)id_row[id] * bs + c];
}

__global__ void k_data_unpack(int64_t nnode, int nvd_file, int bs, const int32_t *__restrict__ id_level, const int32_t *__restrict__ id_row,
                              double *const *__restrict__ vp, const double *__restrict__ body)
{
  const int total = nvd_file * bs;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnode * total) return;
  const int64_t id = i / total;
  const int s = (int)(i - id * total), v = s / bs, c = s - v * bs;
  double *dst = vp[v * UGGPU_MAX_LEVELS + id_level[id]];
  if (dst) dst[(size_t)id_row[id] * bs + c] = body[i];         // entry[j] < 0: value skipped (data_io.cc:596)
}

// checks the node map and collects the device pointers of the vectors on the levels the map names
static int gather_tables(uggpu_ctx *ctx, int nvd, const int *vec, int64_t nnode, const int32_t *id_level, const int32_t *id_row, int *bs_out,
                         std::vector<const double *> &vp)
{
  if (nvd < 1 || nvd > 100 || !vec || nnode < 0 || (nnode && (!id_level || !id_row))) return uggpu_fail(UGGPU_ERROR, "savedata / loaddata: bad argument");
  bool used[UGGPU_MAX_LEVELS] = {false};
  int bs = 0;
  for (int64_t i = 0; i < nnode; i++) {
    const int l = id_level[i];
    if (l < 0 || l >= UGGPU_MAX_LEVELS || !ctx->lev[l].exists || id_row[i] < 0 || id_row[i] >= ctx->lev[l].n)
      return uggpu_fail(UGGPU_ERROR, "savedata / loaddata: node %lld maps to level %d row %d, which does not exist", (long long)i, l, id_row[i]);
    used[l] = true;
  }
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) if (used[l]) {
    if (bs && ctx->lev[l].bs != bs) return uggpu_fail(UGGPU_DESC_MISMATCH, "savedata / loaddata: levels with different numbers of components");
    bs = ctx->lev[l].bs;
    if (ctx->lev[l].partitioned) return uggpu_fail(UGGPU_ERROR, "savedata / loaddata: level %d is partitioned (one file per rank is not offered)", l);
  }
  vp.assign((size_t)nvd * UGGPU_MAX_LEVELS, nullptr);
  for (int v = 0; v < nvd; v++) {
    if (vec[v] < 0) continue;
    for (int l = 0; l < UGGPU_MAX_LEVELS; l++) if (used[l]) {
      const double *p = get_vec(ctx, l, vec[v]);
      if (!p) return UGGPU_DESC_MISMATCH;
      vp[(size_t)v * UGGPU_MAX_LEVELS + l] = p;
    }
  }
  *bs_out = bs ? bs : 1;
  return 0;
}

extern "C" int uggpu_savedata(uggpu_ctx *ctx, const char *filename, const char *type, const uggpu_data_general *g, int nvd, const int *vec,
                              const char *const *vdname, const char *const *compnames, int64_t nnode, const int32_t *id_level, const int32_t *id_row)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  int bs = 1;
  std::vector<const double *> vp;
  UG_TRY(gather_tables(ctx, nvd, vec, nnode, id_level, id_row, &bs, vp));
  for (int v = 0; v < nvd; v++) if (vec[v] < 0) return uggpu_fail(UGGPU_ERROR, "uggpu_savedata: vector %d missing", v);
  const int64_t nval = nnode * nvd * bs;
  std::vector<double> body((size_t)nval);
  cudaStream_t st = ctx->stream;
  int32_t *d_l = nullptr, *d_r = nullptr; const double **d_vp = nullptr; double *d_body = nullptr;
  int rc = 0;
  if (nval > 0) {
    if (!rc) rc = dalloc(ctx, &d_l, (size_t)nnode);
    if (!rc) rc = dalloc(ctx, &d_r, (size_t)nnode);
    if (!rc) rc = dalloc(ctx, &d_vp, vp.size());
    if (!rc) rc = dalloc(ctx, &d_body, (size_t)nval);
    if (!rc) {
      cudaMemcpyAsync(d_l, id_level, sizeof(int32_t) * (size_t)nnode, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(d_r, id_row, sizeof(int32_t) * (size_t)nnode, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(d_vp, vp.data(), sizeof(double *) * vp.size(), cudaMemcpyHostToDevice, st);
      k_data_pack<<<(unsigned)((nval + 255) / 256), 256, 0, st>>>(nnode, nvd, bs, d_l, d_r, d_vp, d_body);
      ctx->launches++;
      cudaMemcpyAsync(body.data(), d_body, sizeof(double) * (size_t)nval, cudaMemcpyDeviceToHost, st);
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "uggpu_savedata: %s", cudaGetErrorString(e));
    }
    if (d_l) dfree(ctx, d_l, (size_t)nnode);
    if (d_r) dfree(ctx, d_r, (size_t)nnode);
    if (d_vp) dfree(ctx, d_vp, vp.size());
    if (d_body) dfree(ctx, d_body, (size_t)nval);
  }
  if (rc) return rc;
  std::vector<int> ncomp((size_t)nvd, bs);
  return uggpu_data_write(filename, type, g, nvd, ncomp.data(), vdname, compnames, nnode, body.data());
}

extern "C" int uggpu_loaddata(uggpu_ctx *ctx, const char *filename, int nvd, const int *vec, int64_t nnode, const int32_t *id_level,
                              const int32_t *id_row, uggpu_data_general *general_out)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  uggpu_data_general g;
  int nvd_file = 0, ncomp[100];
  int64_t ndata = 0;
  UG_TRY(uggpu_data_read(filename, &g, &nvd_file, ncomp, 100, &ndata, nullptr, 0));
  if (general_out) *general_out = g;
  // the caller's list may be shorter than the file's (data_io.cc:515: descriptors beyond n, or NULL ones, are skipped)
  std::vector<int> vf((size_t)nvd_file, -1);
  for (int i = 0; i < nvd_file && i < nvd; i++) vf[i] = vec[i];
  int bs = 1;
  std::vector<const double *> vp;
  UG_TRY(gather_tables(ctx, nvd_file, vf.data(), nnode, id_level, id_row, &bs, vp));
  for (int i = 0; i < nvd_file; i++)
    if (ncomp[i] != bs) return uggpu_fail(UGGPU_DESC_MISMATCH, "uggpu_loaddata: vd-comp do not match (file %d, level %d)", ncomp[i], bs);   // data_io.cc:521
  if (ndata != nnode * nvd_file * bs) return uggpu_fail(UGGPU_ERROR, "uggpu_loaddata: the file holds %lld values, the node map asks for %lld", (long long)ndata, (long long)(nnode * nvd_file * bs));
  if (ndata == 0) return 0;
  std::vector<double> body((size_t)ndata);
  UG_TRY(uggpu_data_read(filename, &g, &nvd_file, ncomp, 100, &ndata, body.data(), ndata));
  cudaStream_t st = ctx->stream;
  int32_t *d_l = nullptr, *d_r = nullptr; const double **d_vp = nullptr; double *d_body = nullptr;
  int rc = 0;
  if (!rc) rc = dalloc(ctx, &d_l, (size_t)nnode);
  if (!rc) rc = dalloc(ctx, &d_r, (size_t)nnode);
  if (!rc) rc = dalloc(ctx, &d_vp, vp.size());
  if (!rc) rc = dalloc(ctx, &d_body, (size_t)ndata);
  if (!rc) {
    cudaMemcpyAsync(d_l, id_level, sizeof(int32_t) * (size_t)nnode, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_r, id_row, sizeof(int32_t) * (size_t)nnode, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_vp, vp.data(), sizeof(double *) * vp.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_body, body.data(), sizeof(double) * (size_t)ndata, cudaMemcpyHostToDevice, st);
    k_data_unpack<<<(unsigned)((ndata + 255) / 256), 256, 0, st>>>(nnode, nvd_file, bs, d_l, d_r, (double *const *)d_vp, d_body);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = uggpu_fail(UGGPU_CUDA_ERROR, "uggpu_loaddata: %s", cudaGetErrorString(e));
  }
  if (d_l) dfree(ctx, d_l, (size_t)nnode);
  if (d_r) dfree(ctx, d_r, (size_t)nnode);
  if (d_vp) dfree(ctx, d_vp, vp.size());
  if (d_body) dfree(ctx, d_body, (size_t)ndata);
  return rc;
}
