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
#include <algorithm>
#include <cstring>
#include <map>
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
#define P2P_FLAG_BYTES 1024               // flag block at the start of every window allocation: per rank [started, landed, window seq, spare] (8 bytes each)
#define P2P_FLAG_WORDS 4

// Three transports for the halo copy of a partitioned level (UGGPU_HALO = ghost | window | nccl; default ghost, falling back to
// nccl when CUDA IPC is unavailable on any rank):
//   ghost   the neighbours' copies of the vector are mapped with CUDA IPC (one handle exchange per (level, vector), on first use) and
//           interface rows are stored straight into their ghost rows: by the epilogue of the kernel that computes them in the fused
//           cycle (HaloK, uggpu_internal.h), by ONE small kernel (k_halo_xchg) otherwise.  No staging, no unpack.
//   window  round 1: one receive window per rank; k_halo_push writes the interface rows into the neighbours' windows, k_halo_wait_unpack
//           copies the window into the ghost rows.  Two kernels per exchange; kept for A/B runs.
//   nccl    pack kernel + ncclSend/ncclRecv group.
enum { MODE_NCCL = 1, MODE_WINDOW = 2, MODE_GHOST = 3 };

struct Peer { unsigned char *base = nullptr; int64_t half = 0; };
struct Opened { int rank; cudaIpcMemHandle_t h; void *ptr; };
struct Comm {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double *sendbuf = nullptr;
  size_t sendbuf_cap = 0;
  int64_t exchanges = 0, allreduces = 0;
  // peer-memory paths
  bool p2p_tried = false;
  int mode = MODE_NCCL;
  unsigned char *win = nullptr; size_t win_bytes = 0; int64_t half = 0;   // my window: flag block, then (window transport) 2 * half doubles
  Peer peers[P2P_MAX_RANKS];
  unsigned long long seq = 0;            // window transport: exchanges
  unsigned long long kseq = 0;           // ghost transport: comm kernels
  unsigned int *push_counter = nullptr;
  unsigned long long *go = nullptr;      // ghost transport: the local go words of the comm kernels (HaloK::go)
  std::vector<Opened> opened;            // vectors of other ranks mapped into this process
  std::vector<std::pair<void *, size_t>> graveyard;   // my vectors that other ranks have mapped: freed with the communicator
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

struct GhostMap { double **d_peer = nullptr; double *h_peer[HALO_MAX_NB]; };
struct LevelHalo {
  HaloDev h_dev;
  HaloDev *d_dev = nullptr;
  uint32_t *snd_bits = nullptr; int32_t *snd_first = nullptr, *srow_ptr = nullptr; uint32_t *snd_ent = nullptr;
  size_t nsl = 0, n_rows = 0, n_ent = 0;
  std::vector<int> peer_recv_off;        // where my rows start in neighbour k's ghost region
  std::map<double *, GhostMap> maps;
};


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

const uint32_t *halo_snd_bits(const Level *L) { return L->halo ? L->halo->snd_bits : nullptr; }

static void mat_comm_free(uggpu_ctx *ctx, SellMat *m)
{
  if (m->comm_flag) dfree(ctx, m->comm_flag, ((size_t)(m->n > 0 ? m->n : 0) + 31) / 32 + 1);
  if (m->x_comm) stx_free(ctx, m);
  trc_free_comm(ctx, m);
}

static void level_halo_free(uggpu_ctx *ctx, Level *L)
{
  // what was derived from the send lists: per-slice flags and exception-row lists of the level's matrices
  for (auto &kv : L->mats) mat_comm_free(ctx, &kv.second);
  mat_comm_free(ctx, &L->P);
  mat_comm_free(ctx, &L->R);
  L->last_pushed = nullptr;
  LevelHalo *H = L->halo;
  if (!H) return;
  for (auto &kv : H->maps) if (kv.second.d_peer) { double **p = kv.second.d_peer; dfree(ctx, p, (size_t)HALO_MAX_NB); }
  if (H->d_dev) dfree(ctx, H->d_dev, 1);
  if (H->snd_bits) dfree(ctx, H->snd_bits, H->nsl + 1);
  if (H->snd_first) dfree(ctx, H->snd_first, H->nsl + 1);
  if (H->srow_ptr) dfree(ctx, H->srow_ptr, H->n_rows + 1);
  if (H->snd_ent) dfree(ctx, H->snd_ent, H->n_ent + 1);
  delete H;
  L->halo = nullptr;
}

// Drops everything that was derived from the set of partitioned levels (window, mapped flag blocks, per-level tables): the next exchange
// sets it up again for the levels that exist then.  Collective in effect: all ranks change their hierarchies together.
static int p2p_reset(uggpu_ctx *ctx, Comm *c)
{
  if (!c->p2p_tried) return 0;
  cudaStreamSynchronize(ctx->stream);
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) level_halo_free(ctx, &ctx->lev[l]);
  for (int q = 0; q < P2P_MAX_RANKS; q++) c->peers[q] = Peer();      // the mappings stay in c->opened until the communicator goes
  if (c->win) { c->graveyard.push_back({c->win, c->win_bytes}); c->win = nullptr; }     // neighbours may still have it mapped
  c->p2p_tried = false;
  c->mode = MODE_NCCL;
  return 0;
}

extern "C" int uggpu_comm_destroy(uggpu_ctx *ctx)
{
  if (!ctx || !ctx->comm) return 0;
  Comm *c = (Comm *)ctx->comm;
  cudaStreamSynchronize(ctx->stream);
  p2p_reset(ctx, c);
  if (c->sendbuf) dfree(ctx, c->sendbuf, c->sendbuf_cap);
  for (auto &o : c->opened) cudaIpcCloseMemHandle(o.ptr);
  for (auto &g : c->graveyard) dev_free(ctx, g.first, g.second);
  if (c->push_counter) dfree(ctx, c->push_counter, 1);
  if (c->go) dfree(ctx, c->go, (size_t)HALO_GO_SLOTS * 16);
  if (c->comm) nccl.CommDestroy(c->comm);
  delete c;
  ctx->comm = nullptr;
  return 0;
}

extern "C" int uggpu_comm_size(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->nranks : 1; }
extern "C" int uggpu_comm_rank(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->rank : 0; }
extern "C" int64_t uggpu_comm_exchanges(uggpu_ctx *ctx) { return ctx && ctx->comm ? ((Comm *)ctx->comm)->exchanges : 0; }
extern "C" int uggpu_comm_transport(uggpu_ctx *ctx)
{
  if (!ctx || !ctx->comm) return UGGPU_TRANSPORT_NONE;
  Comm *c = (Comm *)ctx->comm;
  return c->p2p_tried ? c->mode : UGGPU_TRANSPORT_NONE;
}

int level_free_part(uggpu_ctx *ctx, Level *L)
{
  if (ctx->comm && (L->partitioned || L->halo)) p2p_reset(ctx, (Comm *)ctx->comm);
  if (L->d_part) dfree(ctx, L->d_part, 1);
  if (L->d_send_idx) dfree(ctx, L->d_send_idx, (size_t)L->send_total);
  delete L->part;
  L->part = nullptr;
  L->send_total = 0;
  return 0;
}

// A vector of a partitioned level is released: when other ranks have it mapped (ghost transport) its memory must outlive their
// mappings -- it is parked until the communicator goes.  Returns true when the caller must NOT free it.
bool halo_vec_release(uggpu_ctx *ctx, Level *L, double *p, size_t bytes)
{
  if (!ctx->comm || !L->halo) return false;
  auto it = L->halo->maps.find(p);
  if (it == L->halo->maps.end()) return false;
  if (it->second.d_peer) { double **d = it->second.d_peer; dfree(ctx, d, (size_t)HALO_MAX_NB); }
  L->halo->maps.erase(it);
  if (L->last_pushed == p) L->last_pushed = nullptr;
  ((Comm *)ctx->comm)->graveyard.push_back({p, bytes});
  return true;
}

__global__ void k_halo_pack(int total, int bs, const int32_t *__restrict__ idx, const double *__restrict__ v, double *__restrict__ buf)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total * bs) return;
  int e = i / bs, c = i - e * bs;
  buf[i] = v[(size_t)idx[e] * bs + c];
}

// ---- window transport ------------------------------------------------------------------------------------------------------
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

// ---- ghost transport, the stand-alone exchange ---------------------------------------------------------------------------------
// One kernel: block 0 publishes `started = g`; every block waits until all neighbours have started kernel g (their reads of the ghost
// rows about to be overwritten are done), stores its share of the interface rows into the neighbours' ghost rows; the last block to
// finish publishes `landed = g` and waits for the neighbours' `landed = g`, so the kernel ends when MY ghost rows are complete.
struct XchgArgs {
  int nnb, bs;
  int send_off[HALO_MAX_NB + 1];
  double *dst[HALO_MAX_NB];                  // neighbour k's ghost rows of the vector, at my first row there
};

__global__ void __launch_bounds__(256) k_halo_xchg(XchgArgs a, const HaloDev *__restrict__ dev, unsigned long long g, int total, const int32_t *__restrict__ idx,
                                                   const double *__restrict__ v, unsigned int *counter, int *err)
{
  if (blockIdx.x == 0) halo_publish_dev(dev, g, 0);
  if (threadIdx.x < 32) halo_wait_dev(dev, g, 0, err);
  __syncthreads();
  const int n = total * a.bs;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i / a.bs, cpt = i - e * a.bs;
    int k = 0;
    while (e >= a.send_off[k + 1]) k++;
    a.dst[k][(size_t)(e - a.send_off[k]) * a.bs + cpt] = v[(size_t)idx[e] * a.bs + cpt];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    halo_publish_dev(dev, g, 1);
    if (threadIdx.x < 32) halo_wait_dev(dev, g, 1, err);
    if (threadIdx.x == 0) *counter = 0;
  }
}

// flag[s] bit 0: a row of slice s has an entry in a ghost column (column index >= n_owned); bit 1: a row of slice s is pushed
__global__ void k_comm_flag(SellView A, int n_owned, const uint32_t *__restrict__ snd_bits, uint8_t *__restrict__ flag)
{
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if ((r & ~31) >= A.n) return;
  bool ghost = false;
  if (r < A.n && n_owned >= 0) {
    const int len = A.rowlen[r];
    const ColIter ci = col_iter(A, r);
    for (int j = 0; j < len; j++) if (col_at(ci, j) >= n_owned) ghost = true;
  }
  ghost = __any_sync(0xffffffffu, ghost);
  if ((threadIdx.x & 31) == 0) flag[r >> 5] = (ghost ? 1 : 0) | ((snd_bits && snd_bits[r >> 5]) ? 2 : 0);
}

// offset of a device pointer inside the allocation its IPC handle stands for (cudaMalloc packs small allocations into shared blocks; the
// handle of such a pointer is the block's, and the importer gets the block's base)
typedef int (*GetAddressRangeFn)(unsigned long long *, size_t *, unsigned long long);
static bool ipc_offset(const void *p, int64_t *offset)
{
  static GetAddressRangeFn range = nullptr;
  if (!range) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) range = (GetAddressRangeFn)fn; else cudaGetLastError();
  }
  unsigned long long base = 0; size_t size = 0;
  if (!range || range(&base, &size, (unsigned long long)(uintptr_t)p) != 0) return false;
  *offset = (int64_t)((unsigned long long)(uintptr_t)p - base);
  return true;
}

// A partitioned level takes part in the exchanges even when THIS rank has no neighbour on it (the numbering of the comm kernels and the
// collective set-up steps must stay the same on all ranks)
static bool level_comm(const uggpu_ctx *ctx, const Level *L) { return ctx->comm && L->exists && L->partitioned; }

// Collective, at the first halo operation after the set of partitioned levels changed: chooses the transport, allocates my flag block
// (+ window), exchanges the IPC handles with ncclAllGather, maps the flag blocks of all ranks that are a neighbour on some level.
// All ranks agree on the outcome (all-reduce of the success flags); on failure the NCCL send/recv path stays in use.
static int p2p_setup(uggpu_ctx *ctx, Comm *c)
{
  c->p2p_tried = true;
  const char *mode = getenv("UGGPU_HALO");
  int want = MODE_GHOST;
  if (mode && strcmp(mode, "nccl") == 0) want = MODE_NCCL;
  else if (mode && strcmp(mode, "window") == 0) want = MODE_WINDOW;
  if (c->nranks > P2P_MAX_RANKS) want = MODE_NCCL;
  int64_t half = 0;
  bool nb[P2P_MAX_RANKS] = {false};
  bool any = false;
  for (int l = 0; l < UGGPU_MAX_LEVELS; l++) {
    Level &L = ctx->lev[l];
    if (!level_comm(ctx, &L)) continue;
    any = true;
    for (int k = 0; k < L.nnb; k++) if (L.nb_rank[k] >= 0 && L.nb_rank[k] < P2P_MAX_RANKS) nb[L.nb_rank[k]] = true;
    int64_t h = (int64_t)L.nghost * L.bs;
    if (h > half) half = h;
  }
  half = (half + 31) & ~(int64_t)31;
  if (want != MODE_WINDOW) half = 0;
  struct Info { cudaIpcMemHandle_t h; int64_t half; int64_t ok; int64_t offset; };
  static_assert(sizeof(Info) == 88, "Info layout");
  Info mine;
  memset(&mine, 0, sizeof mine);
  mine.half = half; mine.ok = 0;
  if (want != MODE_NCCL) {
    c->win_bytes = P2P_FLAG_BYTES + 2 * (size_t)half * sizeof(double);
    if (dev_alloc(ctx, (void **)&c->win, c->win_bytes) == 0 && (c->push_counter || dalloc(ctx, &c->push_counter, 1) == 0)) {
      cudaMemsetAsync(c->win, 0, c->win_bytes, ctx->stream);
      cudaMemsetAsync(c->push_counter, 0, sizeof(unsigned int), ctx->stream);
      if (ipc_offset(c->win, &mine.offset) && cudaIpcGetMemHandle(&mine.h, c->win) == cudaSuccess) mine.ok = 1; else cudaGetLastError();
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
  if (ok && any) {
    for (int q = 0; q < c->nranks && ok; q++) {
      if (!nb[q] || q == c->rank) continue;
      void *ptr = nullptr;
      for (auto &o : c->opened) if (o.rank == q && memcmp(&o.h, &all[q].h, sizeof(cudaIpcMemHandle_t)) == 0) ptr = o.ptr;
      if (!ptr) {
        if (cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
        c->opened.push_back(Opened{q, all[q].h, ptr});
      }
      c->peers[q].base = (unsigned char *)ptr + all[q].offset;
      c->peers[q].half = all[q].half;
    }
  }
  // agree: 1.0 per rank that opened everything
  double *d_ok = (double *)d_send, h_ok = ok ? 1.0 : 0.0;
  CUDA_TRY(cudaMemcpyAsync(d_ok, &h_ok, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(nccl.AllReduce(d_ok, d_ok, 1, ncclDouble, ncclSum, c->comm, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(&h_ok, d_ok, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  c->mode = (h_ok == (double)c->nranks) ? want : MODE_NCCL;
  dfree(ctx, d_send, 1); dfree(ctx, d_recv, (size_t)c->nranks);
  if (getenv("UGGPU_HALO_VERBOSE") && c->rank == 0)
    fprintf(stderr, "uggpu: halo exchange over %s (%d ranks, window %lld doubles per half)\n",
            c->mode == MODE_GHOST ? "peer-memory ghost rows (CUDA IPC)" : c->mode == MODE_WINDOW ? "peer-memory windows (CUDA IPC)" : "NCCL send/recv", c->nranks, (long long)half);
  return 0;
}

// Collective, once per partitioned level: every rank publishes its interface description (neighbour ranks, receive offsets, send
// counts); from the neighbours' tables follow the places of my rows in their ghost regions.  Then the device tables of the level:
// the flag words, and the send map (row -> (neighbour, index in its ghost region)) used by the kernels that push.
static int level_halo_setup(uggpu_ctx *ctx, Comm *c, Level *L)
{
  if (L->halo) return 0;
  if (L->nnb > HALO_MAX_NB) return uggpu_fail(UGGPU_ERROR, "halo: %d neighbours on one level (at most %d)", L->nnb, HALO_MAX_NB);
  struct Tab { int32_t nnb, pad; int32_t rank[HALO_MAX_NB], recv_off[HALO_MAX_NB + 1], send_cnt[HALO_MAX_NB]; int32_t fill; };
  Tab mine;
  memset(&mine, 0, sizeof mine);
  mine.nnb = L->nnb;
  for (int k = 0; k < L->nnb; k++) { mine.rank[k] = L->nb_rank[k]; mine.recv_off[k] = L->nb_recv_off[k]; mine.send_cnt[k] = L->nb_send_off[k + 1] - L->nb_send_off[k]; }
  mine.recv_off[L->nnb] = L->nb_recv_off[L->nnb];
  Tab *d_send = nullptr, *d_recv = nullptr;
  UG_TRY(dalloc(ctx, &d_send, 1));
  UG_TRY(dalloc(ctx, &d_recv, (size_t)c->nranks));
  CUDA_TRY(cudaMemcpyAsync(d_send, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(nccl.AllGather(d_send, d_recv, sizeof(Tab), ncclChar, c->comm, ctx->stream));
  std::vector<Tab> all((size_t)c->nranks);
  CUDA_TRY(cudaMemcpyAsync(all.data(), d_recv, sizeof(Tab) * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_send, 1); dfree(ctx, d_recv, (size_t)c->nranks);
  LevelHalo *H = new LevelHalo();
  memset(&H->h_dev, 0, sizeof H->h_dev);
  for (int k = 0; k < L->nnb; k++) {
    const int q = L->nb_rank[k];
    if (q < 0 || q >= c->nranks || q == c->rank) { delete H; return uggpu_fail(UGGPU_ERROR, "halo: bad neighbour rank %d", q); }
    const Tab &t = all[q];
    int kk = -1;
    for (int j = 0; j < t.nnb; j++) if (t.rank[j] == c->rank) kk = j;
    if (kk < 0) { delete H; return uggpu_fail(UGGPU_ERROR, "halo: rank %d does not list rank %d as a neighbour", q, c->rank); }
    if (t.recv_off[kk + 1] - t.recv_off[kk] != mine.send_cnt[k] || t.send_cnt[kk] != mine.recv_off[k + 1] - mine.recv_off[k]) {
      delete H;
      return uggpu_fail(UGGPU_ERROR, "halo: interface sizes of ranks %d and %d differ", c->rank, q);
    }
    H->peer_recv_off.push_back(t.recv_off[kk]);
  }
  L->halo = H;
  L->peer_recv_off = H->peer_recv_off;
  if (c->mode != MODE_GHOST) return 0;
  // flag words
  HaloDev &D = H->h_dev;
  D.nnb = L->nnb;
  for (int k = 0; k < L->nnb; k++) {
    const int q = L->nb_rank[k];
    if (!c->peers[q].base) return uggpu_fail(UGGPU_ERROR, "halo: the flag block of rank %d is not mapped", q);
    D.my_flag[k] = reinterpret_cast<const unsigned long long *>(c->win) + (size_t)q * P2P_FLAG_WORDS;
    D.peer_flag[k] = reinterpret_cast<unsigned long long *>(c->peers[q].base) + (size_t)c->rank * P2P_FLAG_WORDS;
  }
  // send map: rows ascending, per row its (neighbour, destination) pairs
  std::vector<int32_t> idx((size_t)L->send_total);
  if (L->send_total) CUDA_TRY(cudaMemcpyAsync(idx.data(), L->d_send_idx, sizeof(int32_t) * idx.size(), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  std::vector<std::pair<int32_t, uint32_t>> ent;
  ent.reserve(idx.size());
  for (int k = 0; k < L->nnb; k++)
    for (int e = L->nb_send_off[k]; e < L->nb_send_off[k + 1]; e++) {
      const int64_t dst = e - L->nb_send_off[k];      // relative to the place of my rows in neighbour k's ghost region (GhostMap::h_peer[k] points there)
      if (dst >= (1 << 27) || idx[e] < 0 || idx[e] >= L->n) return uggpu_fail(UGGPU_ERROR, "halo: send list entry out of range");
      ent.push_back({idx[e], ((uint32_t)k << 27) | (uint32_t)dst});
    }
  std::stable_sort(ent.begin(), ent.end(), [](const std::pair<int32_t, uint32_t> &a, const std::pair<int32_t, uint32_t> &b) { return a.first < b.first; });
  const size_t nsl = ((size_t)L->n + 31) / 32;
  std::vector<uint32_t> bits(nsl + 1, 0u), ents;
  std::vector<int32_t> first(nsl + 1, 0), rptr;
  ents.reserve(ent.size());
  int32_t rows = 0;
  for (size_t i = 0; i < ent.size();) {
    const int32_t r = ent[i].first;
    bits[(size_t)r >> 5] |= 1u << (r & 31);
    rptr.push_back((int32_t)ents.size());
    for (; i < ent.size() && ent[i].first == r; i++) ents.push_back(ent[i].second);
    rows++;
  }
  rptr.push_back((int32_t)ents.size());
  { int32_t acc = 0; for (size_t s = 0; s < nsl; s++) { first[s] = acc; acc += __builtin_popcount(bits[s]); } first[nsl] = acc; }
  H->nsl = nsl; H->n_rows = (size_t)rows; H->n_ent = ents.size();
  UG_TRY(dalloc(ctx, &H->snd_bits, nsl + 1)); UG_TRY(dalloc(ctx, &H->snd_first, nsl + 1));
  UG_TRY(dalloc(ctx, &H->srow_ptr, H->n_rows + 1)); UG_TRY(dalloc(ctx, &H->snd_ent, H->n_ent + 1));
  cudaStream_t st = ctx->stream;
  CUDA_TRY(cudaMemcpyAsync(H->snd_bits, bits.data(), sizeof(uint32_t) * (nsl + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(H->snd_first, first.data(), sizeof(int32_t) * (nsl + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(H->srow_ptr, rptr.data(), sizeof(int32_t) * rptr.size(), cudaMemcpyHostToDevice, st));
  if (!ents.empty()) CUDA_TRY(cudaMemcpyAsync(H->snd_ent, ents.data(), sizeof(uint32_t) * ents.size(), cudaMemcpyHostToDevice, st));
  D.snd_bits = H->snd_bits; D.snd_first = H->snd_first; D.srow_ptr = H->srow_ptr; D.snd_ent = H->snd_ent;
  UG_TRY(dalloc(ctx, &H->d_dev, 1));
  CUDA_TRY(cudaMemcpyAsync(H->d_dev, &D, sizeof(HaloDev), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

// The neighbours' copies of vector v of level L (ghost transport).  Collective on first use of a (level, vector) pair: all ranks
// exchange the IPC handle of their copy, every rank maps its neighbours'.  h_peer[k] / d_peer[k] = neighbour k's ghost region of the
// vector at the place where MY rows start.
static int ghost_map(uggpu_ctx *ctx, Comm *c, Level *L, double *v, GhostMap **out)
{
  LevelHalo *H = L->halo;
  auto it = H->maps.find(v);
  if (it != H->maps.end()) { *out = &it->second; return 0; }
  struct Info { cudaIpcMemHandle_t h; int64_t offset; int64_t own; int64_t ok; };
  static_assert(sizeof(Info) == 88, "Info layout");
  Info mine;
  memset(&mine, 0, sizeof mine);
  mine.own = (int64_t)L->n * L->bs;
  if (ipc_offset(v, &mine.offset) && cudaIpcGetMemHandle(&mine.h, v) == cudaSuccess) mine.ok = 1; else cudaGetLastError();
  Info *d_send = nullptr, *d_recv = nullptr;
  UG_TRY(dalloc(ctx, &d_send, 1));
  UG_TRY(dalloc(ctx, &d_recv, (size_t)c->nranks));
  CUDA_TRY(cudaMemcpyAsync(d_send, &mine, sizeof mine, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(nccl.AllGather(d_send, d_recv, sizeof(Info), ncclChar, c->comm, ctx->stream));
  std::vector<Info> all((size_t)c->nranks);
  CUDA_TRY(cudaMemcpyAsync(all.data(), d_recv, sizeof(Info) * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  dfree(ctx, d_send, 1); dfree(ctx, d_recv, (size_t)c->nranks);
  for (int q = 0; q < c->nranks; q++) if (!all[q].ok) return uggpu_fail(UGGPU_CUDA_ERROR, "halo: rank %d cannot export a vector of level %d (CUDA IPC)", q, (int)(L - ctx->lev));
  GhostMap gm;
  for (int k = 0; k < L->nnb; k++) {
    const int q = L->nb_rank[k];
    void *ptr = nullptr;
    for (auto &o : c->opened) if (o.rank == q && memcmp(&o.h, &all[q].h, sizeof(cudaIpcMemHandle_t)) == 0) ptr = o.ptr;
    if (!ptr) {
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[q].h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); return uggpu_fail(UGGPU_CUDA_ERROR, "halo: cannot map a vector of rank %d: %s", q, cudaGetErrorString(e)); }
      c->opened.push_back(Opened{q, all[q].h, ptr});
    }
    gm.h_peer[k] = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ptr) + all[q].offset) + all[q].own + (size_t)H->peer_recv_off[k] * L->bs;
  }
  UG_TRY(dalloc(ctx, &gm.d_peer, (size_t)HALO_MAX_NB));
  if (L->nnb) CUDA_TRY(cudaMemcpyAsync(gm.d_peer, gm.h_peer, sizeof(double *) * L->nnb, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  auto ins = H->maps.insert({v, gm});
  *out = &ins.first->second;
  return 0;
}

static int halo_ready(uggpu_ctx *ctx, Comm *c, Level *L)
{
  if (!c->p2p_tried) UG_TRY(p2p_setup(ctx, c));
  if (!L->halo) UG_TRY(level_halo_setup(ctx, c, L));
  return 0;
}

bool halo_fused_available(uggpu_ctx *ctx, int level)
{
  Level *L = &ctx->lev[level];
  if (!level_comm(ctx, L)) return false;
  Comm *c = (Comm *)ctx->comm;
  if (halo_ready(ctx, c, L)) return false;
  return c->mode == MODE_GHOST;
}

int halo_exchange(uggpu_ctx *ctx, int level, double *v);

// slices of M (rows on `RL`, columns index vectors of `CL` or nullptr) that read ghost columns / hold rows to push
static int halo_comm_flag(uggpu_ctx *ctx, Level *RL, Level *CL, SellMat *M)
{
  if (M->comm_flag) return 0;
  const size_t nsl = ((size_t)M->n + 31) / 32;
  UG_TRY(dalloc(ctx, &M->comm_flag, nsl + 1));
  const bool cpart = CL && level_comm(ctx, CL), rpart = level_comm(ctx, RL) && RL->halo && RL->halo->snd_bits && RL->n == M->n;
  if (nsl)
    k_comm_flag<<<(int)((nsl * 32 + 255) / 256), 256, 0, ctx->stream>>>(view(*M), cpart ? CL->n : -1, rpart ? RL->halo->snd_bits : nullptr, M->comm_flag);
  KCHECK(ctx);
  return 0;
}

int halo_prepare(uggpu_ctx *ctx, int row_level, int col_level, SellMat *M, double *operand, const HaloPlan *hp, HaloK *hk)
{
  *hk = halo_none();
  Level *RL = &ctx->lev[row_level], *CL = col_level >= 0 ? &ctx->lev[col_level] : nullptr;
  const bool rpart = level_comm(ctx, RL), cpart = CL && level_comm(ctx, CL);
  if (!rpart && !cpart) return 0;
  Comm *c = (Comm *)ctx->comm;
  if (rpart) UG_TRY(halo_ready(ctx, c, RL));
  if (cpart) UG_TRY(halo_ready(ctx, c, CL));
  if (c->mode != MODE_GHOST || getenv("UGGPU_NO_FUSED_HALO")) {
    if (cpart && operand) return halo_exchange(ctx, col_level, operand);
    return 0;
  }
  const bool ready = hp && hp->operand_ready && cpart && CL->last_pushed == operand;
  if (cpart && operand && !ready) UG_TRY(halo_exchange(ctx, col_level, operand));
  double *push = (hp && hp->push && rpart && hp->push_level == row_level) ? hp->push : nullptr;
  if (!cpart && !push) return 0;                  // nothing to wait for, nothing to push
  UG_TRY(halo_comm_flag(ctx, RL, CL, M));
  hk->flag = M->comm_flag;
  hk->cdev = cpart ? CL->halo->d_dev : nullptr;
  if (push) {
    GhostMap *gm = nullptr;
    UG_TRY(ghost_map(ctx, c, RL, push, &gm));
    hk->pdev = RL->halo->d_dev;
    hk->peer = gm->d_peer;
    RL->last_pushed = push;
    c->exchanges++;
  }
  hk->g = ++c->kseq;
  hk->err = ctx->derr;
  if (!c->go) {
    UG_TRY(dalloc(ctx, &c->go, (size_t)HALO_GO_SLOTS * 16));
    CUDA_TRY(cudaMemsetAsync(c->go, 0, sizeof(unsigned long long) * HALO_GO_SLOTS * 16, ctx->stream));
  }
  hk->go = c->go;
  return 0;
}

// window transport -- push: my interface rows into the neighbours' windows + the release of the exchange number (on the compute stream);
// wait: poll my flag words and copy my window into the ghost rows of v (on `st`, which must be ordered behind the push).
static int halo_p2p(uggpu_ctx *ctx, Comm *c, Level *L, double *v, bool do_push, bool do_wait, cudaStream_t wait_stream)
{
  const int bs = L->bs, nnb = L->nnb;
  const int64_t need = (int64_t)L->nghost * bs;
  if (need > c->half) return uggpu_fail(UGGPU_ERROR, "halo: level needs %lld window doubles, the window holds %lld", (long long)need, (long long)c->half);
  if (do_push) ++c->seq;
  const unsigned long long seq = c->seq;
  const int parity = (int)(seq & 1ull);
  if (do_push) {
    PushArgs pa;
    pa.nnb = nnb; pa.bs = bs; pa.parity = parity; pa.seq = seq;
    for (int k = 0; k <= nnb; k++) pa.send_off[k] = L->nb_send_off[k];
    for (int k = 0; k < nnb; k++) {
      const Peer &p = c->peers[L->nb_rank[k]];
      const int64_t cnt = (int64_t)(L->nb_send_off[k + 1] - L->nb_send_off[k]);
      if (!p.base || ((int64_t)L->peer_recv_off[k] + cnt) * bs > p.half)
        return uggpu_fail(UGGPU_ERROR, "halo: the window of rank %d is not mapped or too small for level %d", L->nb_rank[k], (int)(L - ctx->lev));
      pa.dst[k] = reinterpret_cast<double *>(p.base + P2P_FLAG_BYTES) + (size_t)parity * p.half + (size_t)L->peer_recv_off[k] * bs;
      pa.flag[k] = reinterpret_cast<unsigned long long *>(p.base) + (size_t)c->rank * P2P_FLAG_WORDS + 2;
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
    wa.nnb = nnb; wa.seq = seq;
    for (int k = 0; k < nnb; k++) wa.flag[k] = reinterpret_cast<const unsigned long long *>(c->win) + (size_t)L->nb_rank[k] * P2P_FLAG_WORDS + 2;
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

// Split exchange for kernels that overlap it with their interior rows (spmv.cu k_smooth_step, UGGPU_OVERLAP=1 with the window
// transport): halo_begin pushes on the compute stream and returns 1 when the second half may run on another stream, 0 when there is
// nothing to exchange, and does the whole exchange itself (returning 0) on the other transports.
int halo_begin(uggpu_ctx *ctx, int level, double *v, int *split)
{
  *split = 0;
  Level *L = &ctx->lev[level];
  if (!level_comm(ctx, L)) return 0;
  Comm *c = (Comm *)ctx->comm;
  UG_TRY(halo_ready(ctx, c, L));
  static const bool overlap = getenv("UGGPU_OVERLAP") != nullptr;
  if (c->mode != MODE_WINDOW || !overlap) return halo_exchange(ctx, level, v);
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
  if (!level_comm(ctx, L)) return 0;
  Comm *c = (Comm *)ctx->comm;
  UG_TRY(halo_ready(ctx, c, L));
  const int bs = L->bs;
  ProfScope ps(ctx, UGGPU_K_HALO, level, 16.0 * bs * (double)L->send_total);
  if (c->mode == MODE_GHOST) {
    GhostMap *gm = nullptr;
    UG_TRY(ghost_map(ctx, c, L, v, &gm));
    XchgArgs xa;
    xa.nnb = L->nnb; xa.bs = bs;
    for (int k = 0; k <= L->nnb; k++) xa.send_off[k] = L->nb_send_off[k];
    for (int k = 0; k < L->nnb; k++) xa.dst[k] = gm->h_peer[k];
    const int tot = L->send_total * bs;
    int pb = (tot + 255) / 256;
    if (pb > 4 * ctx->sm_count) pb = 4 * ctx->sm_count;      // every block waits: all of them resident
    if (pb < 1) pb = 1;
    k_halo_xchg<<<pb, 256, 0, ctx->stream>>>(xa, L->halo->d_dev, ++c->kseq, L->send_total, L->d_send_idx, v, c->push_counter, ctx->derr);
    KCHECK(ctx);
    L->last_pushed = nullptr;
    c->exchanges++;
    return 0;
  }
  if (c->mode == MODE_WINDOW) return halo_p2p(ctx, c, L, v, true, true, ctx->stream);
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
  for (int k = 0; k < L->nnb; k++) {
    int ns = L->nb_send_off[k + 1] - L->nb_send_off[k], nr = L->nb_recv_off[k + 1] - L->nb_recv_off[k];
    if (ns > 0) NCCL_TRY(nccl.Send(c->sendbuf + (size_t)L->nb_send_off[k] * bs, (size_t)ns * bs, ncclDouble, L->nb_rank[k], c->comm, ctx->stream));
    if (nr > 0) NCCL_TRY(nccl.Recv(v + ((size_t)L->n + L->nb_recv_off[k]) * bs, (size_t)nr * bs, ncclDouble, L->nb_rank[k], c->comm, ctx->stream));
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
  ProfScope ps(ctx, UGGPU_K_ALLREDUCE, -2, 16.0 * (double)count);
  NCCL_TRY(nccl.AllReduce(dptr, dptr, count, ncclDouble, ncclSum, c->comm, ctx->stream));
  c->allreduces++;
  return 0;
}

// ---- a partition the CALLER supplies (SURVEY.md 8 a13 / e): what a ModelP application knows from DDD ----------------------------------------
// The level was created with n = the rows this rank OWNS (master vectors, parallel/dddif/priority.cc:200-222); its vectors get n_ghost more
// rows at the tail, and column indices >= n in the level's matrices and transfer stencils address them.  Ghost row recv_off[k] + j is
// the copy of the j-th row neighbour nb_rank[k] sends; send_idx[send_off[k] .. send_off[k+1]) are my rows neighbour k needs, in the order
// of ITS ghost rows -- both sides enumerate an interface in the same order (the role of the sorted interface lists of
// parallel/ddd/if/ifcreate.cc:155-203).  Call it after uggpu_level_create and before the first vector of the level exists.  From then on
// every entry point behaves as on the synthetic partitions: halo copies in front of (or fused into) the kernels that read ghost columns,
// global sums in the reductions.  Levels every rank holds completely are simply not given a partition.
extern "C" int uggpu_level_set_partition(uggpu_ctx *ctx, int level, int n_ghost, int64_t n_global, int nnb, const int32_t *nb_rank,
                                         const int32_t *send_off, const int32_t *send_idx, const int32_t *recv_off)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!ctx->comm) return uggpu_fail(UGGPU_ERROR, "uggpu_level_set_partition needs uggpu_comm_init first");
  if (!L->vecs.empty()) return uggpu_fail(UGGPU_ERROR, "level %d already has vectors: set the partition right after uggpu_level_create", level);
  if (n_ghost < 0 || nnb < 0 || nnb > HALO_MAX_NB || (nnb > 0 && (!nb_rank || !send_off || !recv_off)))
    return uggpu_fail(UGGPU_ERROR, "bad partition description (n_ghost %d, %d neighbours, at most %d)", n_ghost, nnb, HALO_MAX_NB);
  Comm *c = (Comm *)ctx->comm;
  if (nnb > 0 && (send_off[0] != 0 || recv_off[0] != 0 || recv_off[nnb] != n_ghost)) return uggpu_fail(UGGPU_ERROR, "partition: offsets must start at 0 and recv_off must end at n_ghost");
  for (int k = 0; k < nnb; k++) {
    if (nb_rank[k] < 0 || nb_rank[k] >= c->nranks || nb_rank[k] == c->rank) return uggpu_fail(UGGPU_ERROR, "partition: bad neighbour rank %d", nb_rank[k]);
    if (send_off[k + 1] < send_off[k] || recv_off[k + 1] < recv_off[k]) return uggpu_fail(UGGPU_ERROR, "partition: offsets must not decrease");
  }
  const int total = nnb > 0 ? send_off[nnb] : 0;
  if (total > 0 && !send_idx) return uggpu_fail(UGGPU_ERROR, "partition: send_idx missing");
  for (int e = 0; e < total; e++) if (send_idx[e] < 0 || send_idx[e] >= L->n) return uggpu_fail(UGGPU_ERROR, "partition: send_idx[%d] = %d is not an owned row", e, send_idx[e]);
  UG_TRY(level_free_part(ctx, L));
  L->partitioned = true;
  L->nghost = n_ghost;
  L->n_global = n_global;
  L->nnb = nnb;
  L->nb_rank.assign(nb_rank, nb_rank + nnb);
  L->nb_send_off.assign(send_off, send_off + nnb + (nnb > 0 ? 1 : 0));
  L->nb_recv_off.assign(recv_off, recv_off + nnb + (nnb > 0 ? 1 : 0));
  if (nnb == 0) { L->nb_send_off.assign(1, 0); L->nb_recv_off.assign(1, 0); }
  L->send_total = total;
  UG_TRY(dalloc(ctx, &L->d_send_idx, (size_t)total));
  if (total > 0) {
    CUDA_TRY(cudaMemcpyAsync(L->d_send_idx, send_idx, sizeof(int32_t) * (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// ---- host-only description of the partition (no GPU needed): used by the CPU tests of the multi-rank logic ------------------
// out[0..5] own lo/hi, out[6] n_own (rows incl. the dummy rows of the padded numbering), out[7] n_ghost, out[8] nnb, out[9..10] pitches; then per neighbour k (stride 16 from out[16]):
// rank, send count, recv count, send box lo/hi (6), recv box lo/hi (6)
extern "C" int uggpu_part_describe(int dim, int cx, int cy, int cz, int px, int py, int pz, int rank, int32_t *out, int cap)
{
  PartGrid g;
  int cells[3] = {cx, cy, cz}, P[3] = {px, py, pz};
  if (part_make(&g, dim, cells, P, rank, 0)) return uggpu_fail(UGGPU_ERROR, "cells %dx%dx%d do not divide over %dx%dx%d ranks", cx, cy, cz, px, py, pz);
  if (cap < 16 + 16 * g.nnb) return uggpu_fail(UGGPU_ERROR, "output too small");
  for (int d = 0; d < 3; d++) { out[d] = g.own.lo[d]; out[3 + d] = g.own.hi[d]; }
  out[6] = g.n_own; out[7] = g.n_ghost; out[8] = g.nnb; out[9] = g.pitch[0]; out[10] = g.pitch[1];
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

// ---- sums over the copies of a vector: l_vector_collect (np/algebra/ugblas.cc:1035), l_vector_consistent (:398) ----------------------
// A ModelP caller that assembles by ELEMENTS holds ADDITIVE vectors: every rank has added the contributions of its own elements to all
// vectors of those elements -- also to the ones another rank is master of, which here are its ghost rows.  collect: the master gets the sum
// over all copies, the other copies 0 (DDD_IFAOneway BorderVectorIF, Gather_VectorComp / Scatter_VectorComp adding, ugblas.cc:334-381);
// consistent: every copy gets the sum (DDD_IFAExchange BorderVectorSymmIF).  With owned rows + ghost rows: every rank sends its ghost
// segments back to their owners (ncclSend/ncclRecv, the reverse of the halo copy), the owner adds what it receives to its rows in the order
// of its neighbour list (ascending rank) -- one addition per copy, deterministic; for vectors with more than two copies the order of the
// additions may differ from DDD's interface order, which the reference does not fix either ("up to summation order") --
// then the ghosts are zeroed (collect) or refreshed from the owners (consistent).  A setup-path operation (right-hand sides, iterates).
__global__ void k_halo_add_back(int cnt, int bs, const int32_t *__restrict__ idx, double *__restrict__ v, const double *__restrict__ buf)
{
  // the copies ONE neighbour sent: its rows are distinct, so every (row, component) has one thread; the neighbours are processed by
  // consecutive launches in list order, which fixes the order of the additions a row with several copies receives
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt * bs) return;
  const int e = i / bs, q = i - e * bs;
  const size_t k = (size_t)idx[e] * bs + q;
  v[k] = v[k] + buf[(size_t)e * bs + q];
}

static int vector_sum_back(uggpu_ctx *ctx, int level, double *v)
{
  Level *L = &ctx->lev[level];
  if (!level_comm(ctx, L)) return 0;
  Comm *c = (Comm *)ctx->comm;
  UG_TRY(halo_ready(ctx, c, L));
  const int bs = L->bs;
  const size_t need = (size_t)L->send_total * bs;
  if (need > c->sendbuf_cap) {
    if (c->sendbuf) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); UG_TRY(dfree(ctx, c->sendbuf, c->sendbuf_cap)); }
    c->sendbuf_cap = need + need / 4;
    UG_TRY(dalloc(ctx, &c->sendbuf, c->sendbuf_cap));
  }
  ProfScope ps(ctx, UGGPU_K_HALO, level, 16.0 * bs * (double)L->send_total);
  NCCL_TRY(nccl.GroupStart());
  for (int k = 0; k < L->nnb; k++) {
    const int ns = L->nb_send_off[k + 1] - L->nb_send_off[k], nr = L->nb_recv_off[k + 1] - L->nb_recv_off[k];
    // reverse direction: my ghost segment of neighbour k goes to k, k's copies of my rows arrive in the order of my send list for k
    if (nr > 0) NCCL_TRY(nccl.Send(v + ((size_t)L->n + L->nb_recv_off[k]) * bs, (size_t)nr * bs, ncclDouble, L->nb_rank[k], c->comm, ctx->stream));
    if (ns > 0) NCCL_TRY(nccl.Recv(c->sendbuf + (size_t)L->nb_send_off[k] * bs, (size_t)ns * bs, ncclDouble, L->nb_rank[k], c->comm, ctx->stream));
  }
  NCCL_TRY(nccl.GroupEnd());
  for (int k = 0; k < L->nnb; k++) {
    const int off = L->nb_send_off[k], cnt = L->nb_send_off[k + 1] - off;
    if (cnt <= 0) continue;
    k_halo_add_back<<<(cnt * bs + 255) / 256, 256, 0, ctx->stream>>>(cnt, bs, L->d_send_idx + off, v, c->sendbuf + (size_t)off * bs);
    KCHECK(ctx);
  }
  c->exchanges++;
  return 0;
}

extern "C" int uggpu_l_vector_collect(uggpu_ctx *ctx, int level, int x)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  double *xp = get_vec(ctx, level, x);
  if (!xp) return UGGPU_DESC_MISMATCH;
  if (!level_comm(ctx, L)) return 0;
  UG_TRY(vector_sum_back(ctx, level, xp));
  if (L->nghost > 0) CUDA_TRY(cudaMemsetAsync(xp + (size_t)L->n * L->bs, 0, sizeof(double) * (size_t)L->nghost * L->bs, ctx->stream));
  L->last_pushed = nullptr;
  return 0;
}

extern "C" int uggpu_l_vector_consistent(uggpu_ctx *ctx, int level, int x)
{
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  double *xp = get_vec(ctx, level, x);
  if (!xp) return UGGPU_DESC_MISMATCH;
  if (!level_comm(ctx, L)) return 0;
  UG_TRY(vector_sum_back(ctx, level, xp));
  return halo_exchange(ctx, level, xp);
}
