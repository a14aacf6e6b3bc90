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
#include <cstdio>
#include <cstdlib>
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
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
} nccl;

#define P2P_MAX_RANKS 32
#define P2P_FLAG_BYTES 512                // flags[rank] (8 bytes each) at the start of every window allocation

// One receive window per rank, opened by its neighbours with CUDA IPC: the halo exchange of a partitioned level is then
//   k_halo_push         my interface rows -> straight into the neighbours' windows over NVLink (peer stores), then one
//                       release store of the exchange number into each neighbour's flag word;
//   k_halo_wait_unpack  wait until every neighbour's number has arrived, copy the window into the ghost rows of the vector.
// No NCCL call, no staging on the sender.  Two window halves alternate: a sender can only be one exchange ahead of a
// receiver (it needs the receiver's flag of exchange e before it starts e+1), so half (e & 1) is free again at e+2.
struct Peer { unsigned char *base = nullptr; int64_t half = 0; };
struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double *sendbuf = nullptr;
  size_t sendbuf_cap = 0;
  int64_t exchanges = 0, allreduces = 0;
  // peer-memory path
  bool p2p_tried = false, p2p = false;
  unsigned char *win = nullptr; size_t win_bytes = 0; int64_t half = 0;   // my window: flags, then 2 * half doubles
  Peer peers[P2P_MAX_RANKS];
  unsigned long long seq = 0;
  unsigned int *push_counter = nullptr;
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
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(Send); SYM(Recv); SYM(AllReduce); SYM(AllGather); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
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
  for (int q = 0; q < P2P_MAX_RANKS; q++) if (c->peers[q].base) cudaIpcCloseMemHandle(c->peers[q].base);
  if (c->win) dfree(ctx, c->win, c->win_bytes);
  if (c->push_counter) dfree(ctx, c->push_counter, 1);
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

// ---- peer-memory halo exchange ---------------------------------------------------------------------------------------------
struct PushArgs {
  int nnb, bs, parity;
  unsigned long long seq;
  int send_off[PART_MAX_NB + 1];
  double *dst[PART_MAX_NB];                  // neighbour's window half + its receive offset for me
  unsigned long long *flag[PART_MAX_NB];     // neighbour's flag word for me
};

__global__ void k_halo_push(PushArgs a, int total, const int32_t *__restrict__ idx, const double *__restrict__ v, unsigned int *counter)
{
  const int n = total * a.bs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i / a.bs, cpt = i - e * a.bs;
    int k = 0;
    while (e >= a.send_off[k + 1]) k++;
    a.dst[k][(size_t)(e - a.send_off[k]) * a.bs + cpt] = v[(size_t)idx[e] * a.bs + cpt];
  }
  // the last block to finish publishes the exchange number to every neighbour (release, system scope)
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < a.nnb) {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[threadIdx.x]), "l"(a.seq) : "memory");
    }
    if (threadIdx.x == 0) *counter = 0;
  }
}

struct WaitArgs {
  int nnb;
  unsigned long long seq;
  const unsigned long long *flag[PART_MAX_NB];   // my flag words, one per neighbour
};

// every block waits for all neighbours (thread k polls neighbour k), then the grid copies the window half into the ghost rows;
// a neighbour that does not show up within ~10 s is reported through the device error word instead of hanging the GPU
__global__ void k_halo_wait_unpack(WaitArgs a, const double *__restrict__ win, double *__restrict__ ghost, size_t count, int *err)
{
  if (threadIdx.x < a.nnb) {
    unsigned long long t0, t1, got;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(a.flag[threadIdx.x]) : "memory");
      if (got >= a.seq) break;
      if (*reinterpret_cast<volatile int *>(err)) break;        // an earlier exchange already failed: do not wait again
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { atomicExch(err, UGGPU_CUDA_ERROR); break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) ghost[i] = win[i];
}

// Allocates my window, exchanges the IPC handles with ncclAllGather, opens the neighbours' windows.  Collective: every rank
// calls it at its first halo exchange.  All ranks agree on the outcome (all-reduce of the success flags); on failure the
// NCCL send/recv path stays in use.
static int p2p_setup(uggpu_ctx *ctx, Comm *c)
{
  c->p2p_tried = true;
  const char *mode = getenv("UGGPU_HALO");
  int want = !(mode && strcmp(mode, "nccl") == 0) && c->nranks <= P2P_MAX_RANKS;
  int64_t half = 0;
  const PartGrid *gany = nullptr;
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) {
    Level &L = ctx->lev[l];
    if (!L.exists || !L.partitioned || !L.part) continue;
    gany = L.part;
    int64_t h = (int64_t)L.nghost * L.bs;
    if (h > half) half = h;
  }
  half = (half + 31) & ~(int64_t)31;
  struct Info { cudaIpcMemHandle_t h; int64_t half; int64_t ok; };
  static_assert(sizeof(Info) == 80, "Info layout");
  Info mine;
  memset(&mine, 0, sizeof mine);
  mine.half = half; mine.ok = 0;
  if (want && gany) {
    c->win_bytes = P2P_FLAG_BYTES + 2 * (size_t)half * sizeof(double);
    if (dev_alloc(ctx, (void **)&c->win, c->win_bytes) == 0 && dalloc(ctx, &c->push_counter, 1) == 0) {
      cudaMemsetAsync(c->win, 0, c->win_bytes, ctx->stream);
      cudaMemsetAsync(c->push_counter, 0, sizeof(unsigned int), ctx->stream);
      if (cudaIpcGetMemHandle(&mine.h, c->win) == cudaSuccess) mine.ok = 1; else cudaGetLastError();
    }
  }
  c->half = half;
  Info *d_send = nullptr, *d_recv = nullptr;
  UG_TRY(dalloc(ctx, &d_send, 1));
  UG_TRY(dalloc(ctx, &d_recv, (size_t)c->nranks));
  CUDA_TRY(cudaMemcpyAsync(d_send, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(nccl.AllGather(d_send, d_recv, sizeof(Info), ncclChar, c->comm, ctx->stream));
  std::vector<Info> all((size_t)c->nranks);
  CUDA_TRY(cudaMemcpyAsync(all.data(), d_recv, sizeof(Info) * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int ok = mine.ok ? 1 : 0;
  for (int q = 0; q < c->nranks; q++) if (!all[q].ok) ok = 0;
  if (ok && gany) {
    for (int k = 0; k < gany->nnb && ok; k++) {
      const int q = gany->nb_rank[k];
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      c->peers[q].base = (unsigned char *)ptr;
      c->peers[q].half = all[q].half;
    }
  }
  // agree: 1.0 per rank that opened everything
  double *d_ok = (double *)d_send, h_ok = ok ? 1.0 : 0.0;
  CUDA_TRY(cudaMemcpyAsync(d_ok, &h_ok, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(nccl.AllReduce(d_ok, d_ok, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&h_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  c->p2p = h_ok == (double)c->nranks;
  dfree(ctx, d_send, 1); dfree(ctx, d_recv, (size_t)c->nranks);
  if (getenv("UGGPU_HALO_VERBOSE") && c->rank == 0)
    fprintf(stderr, "uggpu: halo exchange over %s (%d ranks, window %lld doubles per half)\n", c->p2p ? "peer memory (CUDA IPC)" : "NCCL send/recv", c->nranks, (long long)half);
  return 0;
}

int halo_exchange(uggpu_ctx *ctx, int level, double *v);

// push: my interface rows into the neighbours' windows + the release of the exchange number (on the compute stream);
// wait: poll my flag words and copy my window into the ghost rows of v (on `st`, which must be ordered behind the push).
static int halo_p2p(uggpu_ctx *ctx, Comm *c, Level *L, double *v, bool do_push, bool do_wait, cudaStream_t wait_stream)
{
  const PartGrid &g = *L->part;
  const int bs = L->bs;
  if (L->peer_recv_off.empty()) {       // where my rows land in each neighbour's window: its receive offset for me on this level
    int cells[3];
    for (int d = 0; d < 3; d++) cells[d] = d < g.dim ? g.nn[d] - 1 : 0;
    for (int k = 0; k < g.nnb; k++) {
      PartGrid nb;
      if (part_make(&nb, g.dim, cells, g.P, g.nb_rank[k], 0)) return uggpu_fail(UGGPU_ERROR, "halo: cannot rebuild the partition of rank %d", g.nb_rank[k]);
      int kk = -1;
      for (int j = 0; j < nb.nnb; j++) if (nb.nb_rank[j] == g.rank) kk = j;
      if (kk < 0) return uggpu_fail(UGGPU_ERROR, "halo: rank %d does not list rank %d as a neighbour", g.nb_rank[k], g.rank);
      if (nb.nb_recv_off[kk + 1] - nb.nb_recv_off[kk] != g.nb_send_off[k + 1] - g.nb_send_off[k])
        return uggpu_fail(UGGPU_ERROR, "halo: interface sizes of ranks %d and %d differ", g.rank, g.nb_rank[k]);
      L->peer_recv_off.push_back(nb.nb_recv_off[kk]);
    }
  }
  if (do_push) ++c->seq;
  const unsigned long long seq = c->seq;
  const int parity = (int)(seq & 1ull);
  if (do_push) {
    PushArgs pa;
    pa.nnb = g.nnb; pa.bs = bs; pa.parity = parity; pa.seq = seq;
    for (int k = 0; k <= g.nnb; k++) pa.send_off[k] = g.nb_send_off[k];
    for (int k = 0; k < g.nnb; k++) {
      const Peer &p = c->peers[g.nb_rank[k]];
      pa.dst[k] = reinterpret_cast<double *>(p.base + P2P_FLAG_BYTES) + (size_t)parity * p.half + (size_t)L->peer_recv_off[k] * bs;
      pa.flag[k] = reinterpret_cast<unsigned long long *>(p.base) + c->rank;
    }
    const int tot = L->send_total * bs;
    int pb = (tot + 255) / 256;
    if (pb > 4 * ctx->sm_count) pb = 4 * ctx->sm_count;
    if (pb < 1) pb = 1;
    k_halo_push<<<pb, 256, 0, ctx->stream>>>(pa, L->send_total, L->d_send_idx, v, c->push_counter);
    KCHECK(ctx);
  }
  if (do_wait) {
    WaitArgs wa;
    wa.nnb = g.nnb; wa.seq = seq;
    for (int k = 0; k < g.nnb; k++) wa.flag[k] = reinterpret_cast<const unsigned long long *>(c->win) + g.nb_rank[k];
    const size_t cnt = (size_t)L->nghost * bs;
    int wb = (int)((cnt + 255) / 256);
    if (wb > ctx->sm_count) wb = ctx->sm_count;        // every block polls: keep them all resident
    if (wb < 1) wb = 1;
    k_halo_wait_unpack<<<wb, 256, 0, wait_stream>>>(wa, reinterpret_cast<const double *>(c->win + P2P_FLAG_BYTES) + (size_t)parity * c->half,
                                                     v + (size_t)L->n * bs, cnt, ctx->derr);
    KCHECK(ctx);
    c->exchanges++;
  }
  return 0;
}

static int halo_exchange_p2p(uggpu_ctx *ctx, Comm *c, Level *L, double *v) { return halo_p2p(ctx, c, L, v, true, true, ctx->stream); }

// Split exchange for kernels that overlap it with their interior rows (spmv.cu k_smooth_step): halo_begin pushes on the compute
// stream and returns 1 when the second half may run on another stream (peer-memory path on a partitioned level), 0 when there is
// nothing to exchange, and does the whole exchange itself (returning 0) on the NCCL path.  halo_finish waits + unpacks on `st`;
// the caller orders `st` behind the push (event) and the compute stream behind `st` afterwards.
int halo_begin(uggpu_ctx *ctx, int level, double *v, int *split)
{
  *split = 0;
  Level *L = &ctx->lev[level];
  if (!ctx->comm || !L->partitioned || !L->part || L->part->nnb == 0) return 0;
  Comm *c = (Comm *)ctx->comm;
  if (!c->p2p_tried) UG_TRY(p2p_setup(ctx, c));
  // opt-in (UGGPU_OVERLAP=1): measured SLOWER than the plain exchange on the weak-scaling bench (2 GPUs, 513^3 per GPU: 33.1 vs
  // 29.6 ms per cycle) -- an exchange costs ~25 us next to a 4 ms kernel, while the interface slices of an x-split (every 16th
  // slice of the lexicographic order) run far below the bandwidth of the contiguous sweep.  Kept parity-tested for partitions
  // whose interfaces are contiguous in the row order.
  static const bool overlap = getenv("UGGPU_OVERLAP") != nullptr;
  if (!c->p2p || !overlap) return halo_exchange(ctx, level, v);
  UG_TRY(halo_p2p(ctx, c, L, v, true, false, ctx->stream));
  *split = 1;
  return 0;
}

int halo_finish(uggpu_ctx *ctx, int level, double *v, cudaStream_t st)
{
  Level *L = &ctx->lev[level];
  Comm *c = (Comm *)ctx->comm;
  return halo_p2p(ctx, c, L, v, false, true, st);
}

int halo_exchange(uggpu_ctx *ctx, int level, double *v)
{
  Level *L = &ctx->lev[level];
  if (!ctx->comm || !L->partitioned || !L->part || L->part->nnb == 0) return 0;
  Comm *c = (Comm *)ctx->comm;
  if (!c->p2p_tried) UG_TRY(p2p_setup(ctx, c));
  if (c->p2p) return halo_exchange_p2p(ctx, c, L, v);
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

// ModelP vector consistency.  With owner-computes storage every vector entry is held by exactly one rank, so the
// sums over border copies of l_vector_consistent (np/algebra/ugblas.cc:398) and l_vector_collect (:1035) have nothing to add up;
// what remains of the protocol is the copy of the owners' values into the other ranks' ghost copies, l_ghostvector_consistent
// (:740).  Every entry point does it itself before a kernel reads ghost columns; this is the same exchange for callers that want
// it explicitly (e.g. after uggpu_vec_upload of an iterate).  No-op on one GPU and on levels every rank holds completely.
extern "C" int uggpu_l_ghostvector_consistent(uggpu_ctx *ctx, int level, int x)
{
  if (!get_level(ctx, level)) return UGGPU_ERROR;
  double *xp = get_vec(ctx, level, x);
  if (!xp) return UGGPU_DESC_MISMATCH;
  return halo_exchange(ctx, level, xp);
}
