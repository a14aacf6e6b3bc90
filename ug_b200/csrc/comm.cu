// comm.cu -- NCCL over NVLink for the partitioned levels: replaces, for the hot path only, the reference's
// parallel/ppif + parallel/ddd/if + parallel/dddif call sites listed in SURVEY.md 2.2:
//   DDD_IFAExchange(BorderVectorSymmIF, ...)  in l_vector_consistent (np/algebra/ugblas.cc:398)  -> halo_exchange
//     (a COPY of owner values into the neighbours' ghost rows, see part.h, instead of a SUM over copies);
//   UG_GlobalSumNDOUBLE (parallel/dddif/support.cc:526, binary tree + broadcast)                  -> allreduce_sum
//     (ncclAllReduce on the compute stream);
//   agglomeration of coarse levels (np/procs/amgtransfer.cc:246 $aggLimit, parallel/dddif/lb.cc:160) -> coarse levels
//     below a size threshold are held completely by every rank; the restriction into the first such level is an
//     all-reduce of vectors with disjoint support (= all-gather), the prolongation out of it needs no message.
// NCCL is bound at run time (dlopen) so that single-GPU users of libuggpu.so do not need it.
#include "uggpu_internal.h"
#include "part.h"

#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <vector>

namespace {
struct Nccl {
  void *dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} nccl;

struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double *sendbuf = nullptr;
  size_t sendbuf_cap = 0;
  int64_t exchanges = 0, allreduces = 0;
};

int load_nccl()
{
  if (nccl.dl) return 0;
  const char *names[] = {getenv("UGGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    if (!n || !*n) continue;
    nccl.dl = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (nccl.dl) break;
  }
  if (!nccl.dl) return uggpu_fail(UGGPU_ERROR, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define SYM(f) *(void **)(&nccl.f) = dlsym(nccl.dl, "nccl" #f); if (!nccl.f) return uggpu_fail(UGGPU_ERROR, "NCCL symbol nccl" #f " missing")
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(Send); SYM(Recv); SYM(AllReduce); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
  return 0;
}
}  // namespace

#define NCCL_TRY(expr)                                                                                                       \
  do {                                                                                                                       \
    ncclResult_t r__ = (expr);                                                                                               \
    if (r__ != ncclSuccess) return uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, nccl.GetErrorString(r__)); \
  } while (0)

extern "C" int uggpu_comm_unique_id(void *out128)
{
  UG_TRY(load_nccl());
  ncclUniqueId id;
  NCCL_TRY(nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return 0;
}

extern "C" int uggpu_comm_init(uggpu_ctx *ctx, int nranks, int rank, const void *id128)
{
  if (!ctx || !id128) return uggpu_fail(UGGPU_ERROR, "null argument");
  if (ctx->comm) return uggpu_fail(UGGPU_ERROR, "communicator already initialised");
  UG_TRY(load_nccl());
  CUDA_TRY(cudaSetDevice(ctx->device));
  Comm *c = new Comm();
  c->nranks = nranks; c->rank = rank;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCCL_TRY(nccl.CommInitRank(&c->comm, nranks, id, rank));
  ctx->comm = c;
  return 0;
}

extern "C" int uggpu_comm_destroy(uggpu_ctx *ctx)
{
  if (!ctx || !ctx->comm) return 0;
  Comm *c = (Comm *)ctx->comm;
  cudaStreamSynchronize(ctx->stream);
  if (c->sendbuf) dfree(ctx, c->sendbuf, c->sendbuf_cap);
  if (c->comm) nccl.CommDestroy(c->comm);
  delete c;
  ctx->comm = nullptr;
  return 0;
}

extern "C" int uggpu_comm_size(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->nranks : 1; }
extern "C" int uggpu_comm_rank(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->rank : 0; }
extern "C" int64_t uggpu_comm_exchanges(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->exchanges : 0; }

int level_free_part(uggpu_ctx *ctx, Level *L)
{
  if (L->d_part) dfree(ctx, L->d_part, 1);
  if (L->d_send_idx) dfree(ctx, L->d_send_idx, (size_t)L->send_total);
  delete L->part;
  L->part = nullptr;
  L->send_total = 0;
  return 0;
}

__global__ void k_halo_pack(int total, int bs, const int32_t *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total * bs) return;
  int e = i / bs, c = i - e * bs;
  buf[i] = v[(size_t)idx[e] * bs + c];
}

int halo_exchange(uggpu_ctx *ctx, int level, double *v)
{
  Level *L = &ctx->lev[level];
  if (!ctx->comm || !L->partitioned || !L->part || L->part->nnb == 0) return 0;
  Comm *c = (Comm *)ctx->comm;
  const PartGrid &g = *L->part;
  const int bs = L->bs;
  size_t need = (size_t)L->send_total * bs;
  if (need > c->sendbuf_cap) {
    if (c->sendbuf) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); UG_TRY(dfree(ctx, c->sendbuf, c->sendbuf_cap)); }
    c->sendbuf_cap = need + need / 4;
    UG_TRY(dalloc(ctx, &c->sendbuf, c->sendbuf_cap));
  }
  if (L->send_total > 0) {
    int tot = L->send_total * bs;
    k_halo_pack<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(L->send_total, bs, L->d_send_idx, v, c->sendbuf);
    KCHECK(ctx);
  }
  NCCL_TRY(nccl.GroupStart());
  for (int k = 0; k < g.nnb; k++) {
    int ns = g.nb_send_off[k + 1] - g.nb_send_off[k], nr = g.nb_recv_off[k + 1] - g.nb_recv_off[k];
    if (ns > 0) NCCL_TRY(nccl.Send(c->sendbuf + (size_t)g.nb_send_off[k] * bs, (size_t)ns * bs, ncclDouble, g.nb_rank[k], c->comm, ctx->stream));
    if (nr > 0) NCCL_TRY(nccl.Recv(v + ((size_t)L->n + g.nb_recv_off[k]) * bs, (size_t)nr * bs, ncclDouble, g.nb_rank[k], c->comm, ctx->stream));
  }
  NCCL_TRY(nccl.GroupEnd());
  c->exchanges++;
  return 0;
}

int allreduce_sum(uggpu_ctx *ctx, double *dptr, size_t count)
{
  if (!ctx->comm) return 0;
  Comm *c = (Comm *)ctx->comm;
  if (c->nranks == 1 || count == 0) return 0;
  NCCL_TRY(nccl.AllReduce(dptr, dptr, count, ncclDouble, ncclSum, c->comm, ctx->stream));
  c->allreduces++;
  return 0;
}

// ---- host-only description of the partition (no GPU needed): used by the CPU tests of the multi-rank logic ------------------
// out[0..5] own lo/hi, out[6] n_own, out[7] n_ghost, out[8] nnb; then per neighbour k (stride 16 from out[16]):
// rank, send count, recv count, send box lo/hi (6), recv box lo/hi (6)
extern "C" int uggpu_part_describe(int dim, int cx, int cy, int cz, int px, int py, int pz, int rank, int32_t *out, int cap)
{
  PartGrid g;
  int cells[3] = {cx, cy, cz}, P[3] = {px, py, pz};
  if (part_make(&g, dim, cells, P, rank, 0)) return uggpu_fail(UGGPU_ERROR, "cells %dx%dx%d do not divide over %dx%dx%d ranks", cx, cy, cz, px, py, pz);
  if (cap < 16 + 16 * g.nnb) return uggpu_fail(UGGPU_ERROR, "output too small");
  for (int d = 0; d < 3; d++) { out[d] = g.own.lo[d]; out[3 + d] = g.own.hi[d]; }
  out[6] = g.n_own; out[7] = g.n_ghost; out[8] = g.nnb;
  for (int k = 0; k < g.nnb; k++) {
    int32_t *o = out + 16 + 16 * k;
    o[0] = g.nb_rank[k]; o[1] = g.nb_send_off[k + 1] - g.nb_send_off[k]; o[2] = g.nb_recv_off[k + 1] - g.nb_recv_off[k];
    for (int d = 0; d < 3; d++) { o[3 + d] = g.nb_send[k].lo[d]; o[6 + d] = g.nb_send[k].hi[d]; o[9 + d] = g.nb_recv[k].lo[d]; o[12 + d] = g.nb_recv[k].hi[d]; }
  }
  return 0;
}

// local index (owned or ghost) of global node (x,y,z) for `rank`, or -1: the numbering both the generator and the
// halo lists use (host-only)
extern "C" int uggpu_part_local_index(int dim, int cx, int cy, int cz, int px, int py, int pz, int rank, int x, int y, int z)
{
  PartGrid g;
  int cells[3] = {cx, cy, cz}, P[3] = {px, py, pz};
  if (part_make(&g, dim, cells, P, rank, 0)) return -2;
  int xx[3] = {x, y, z};
  if (!box_has(g.ext, xx)) return -1;
  return part_local_index(g, xx);
}
