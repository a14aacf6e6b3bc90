// uggpu_internal.h -- shared declarations of libuggpu.so (sm_100a device layer of the gpuls numprocs).
//
// Device data layout (DESIGN.md "data layout in HBM"):
//  * vectors: dense double[n*bs], entry (row r, component i) at r*bs+i; row r = position of the UG VECTOR
//    in FIRSTVECTOR->SUCCVC order.
//  * matrices (A per MATDATA_DESC, and the transfer stencils P and R): SELL-32 -- rows are cut into slices of
//    32 consecutive rows (one warp), a slice stores its entries "column-major": entry j of lane l at
//    slice_ptr[s] + j*32 + l.  Entry order inside a row is the canonical one (VSTART->MNEXT, diagonal
//    first), so one thread walking j = 0..len-1 performs the reference's additions in the reference's
//    order while the warp's loads are perfectly coalesced.  Blocks (bs>1) are stored component-planar:
//    component k of entry e at val[(slice_ptr[s] + j*32)*bb + k*32 + l].
//  * column indices are stored per slice in one of two lossless forms (sell.cu sell_compress_cols):
//      explicit  col[col_ptr[s] + j*32 + l]                          (4 B per stored entry)
//      uniform   col[~col_ptr[s] + j] + row                          (4 B per slice column)
//    A slice is "uniform" when entry j of all its rows has the same distance to the row's own index -- the case
//    of every slice of interior rows of a structured or uniformly refined grid.  col_ptr[s] < 0 marks the uniform
//    form (bitwise complement of the offset).  The decoded indices are identical, only fewer bytes cross HBM.
//  * values of uniform slices can be shared the same way (sell.cu sell_share_values, matrices only): when, in addition, entry j
//    of all rows of the slice holds the same value (block) -- every slice of interior rows of a constant-coefficient operator
//    on a structured or uniformly refined grid -- the slice reads its values from a table vt[] (one block per slice column)
//    that slices with bit-identical value vectors share.  The code word of a uniform slice packs both offsets:
//      ~col_ptr[s] = column-table offset (low 32 bits) | (value-table offset + 1) << 32 (0: explicit values)
//    The explicit val[] array stays complete (the ILU/LU setup, uggpu_mat_get and the schedule builders read it); only the
//    SpMV-type kernels take the table, i.e. the bytes that cross HBM per sweep shrink, not the footprint.  Lossless: the
//    decoded doubles are the stored ones bit for bit, verified on the device when the tables are built.
#ifndef UGGPU_INTERNAL_H
#define UGGPU_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

#include "uggpu.h"

#define SLICE 32
// decoding of a uniform slice's code word cp = col_ptr[s] < 0
#define UG_COLTAB(cp) ((int64_t)((~(cp)) & 0xffffffffll))          // offset of the distance table in col[]
#define UG_VALTAB(cp) ((int64_t)((~(cp)) >> 32) - 1)               // offset of the shared value table in vt[]; -1: explicit values

// The dominant stencil of a scalar matrix: the (column distances, values) pair that most slices share (sell_share_values).
// Passed to the stencil variant of the smoothing kernel BY VALUE: the tables then live in the kernel's constant bank and, with
// the loop over the slice columns unrolled, every distance and value is an immediate operand -- no load, no shuffle.
struct Sten {
  long long code;          // col_ptr[s] of the slices that use it
  int w;                   // slice columns (0: the matrix has no dominant stencil)
  int maxd;                // largest column distance (rows first touched through it: prefetch target)
  long long dbytes[32];    // column distance * sizeof(double)
  double v[32];
};

// ... and of a matrix of 3x3 blocks (Q1 hexahedra: 27 block columns): 27 distances + 243 values = 2.2 KB of kernel parameters
struct Sten3 {
  long long code;
  int w, maxd;
  long long dbytes[27];    // column distance * 3 * sizeof(double)
  double v[27 * 9];        // row-major 3x3 blocks in slice-column order
};

struct SellMat {
  int      n = 0;          // rows
  int      bb = 1;         // doubles per entry (bs*bs for A, 1 for transfer weights)
  int64_t  nnz = 0;        // true entries
  int64_t  padded = 0;     // stored entries (multiple of 32 per slice)
  int      maxlen = 0;
  int      fixed_w = 0;    // > 0: every slice is padded to this width (slice_ptr[s] == s*32*fixed_w) and kernels compute the offset instead of loading it
  int64_t *slice_ptr = nullptr;   // [nslices+1], entry offsets
  uint16_t *rowlen = nullptr;     // [n]
  int64_t *col_ptr = nullptr;     // [nslices]; == slice_ptr while all slices are explicit (then it is not a separate allocation)
  int32_t *col = nullptr;         // [col_len]
  int64_t  col_len = 0;           // int32 words in col (== padded while uncompressed)
  int64_t  uniform_slices = 0;
  // transfer stencils only (sell_compress_values): when the matrix holds at most 256 distinct values (the weights of a geometric
  // transfer: 1, 1/2, 1/4, 1/8), one byte per stored entry indexes a table -- the decoded doubles are the stored ones, bit for bit
  uint8_t *vcode = nullptr;       // [padded]
  double  *vtable = nullptr;      // [256]
  int      nvals = 0;
  double  *val = nullptr;         // [padded*bb]
  double  *vt = nullptr;          // [vt_len] shared value tables of uniform slices (sell_share_values), else nullptr
  int64_t  vt_len = 0;
  int64_t  vshared_slices = 0;
  Sten     sten = {0, 0, 0, {0}, {0.0}};   // scalar matrices: the dominant stencil, when more than half of the slices use it
  int64_t  sten_slices = 0;
  Sten3   *sten3 = nullptr;       // 3x3-block matrices: the dominant stencil (host heap, owned by the matrix), else nullptr
  int64_t  val_entries = -1;      // entries whose values a pass fetches from HBM: true entries of explicit slices + the distinct tables (-1: nnz)
  double  *diag = nullptr;        // [nslices*32*bb] copy of entry 0 of every row (the diagonal block), same planar slice layout as val with width 1:
                                  // component k of row r at diag[((r>>5)*bb + k)*32 + (r&31)].  Kernels that need only Diag(A) (l_jac, the Jacobi start
                                  // fused into the restriction) stream it instead of touching one column of every slice of val.  Matrices only (not P, R).
  bool valid() const { return n > 0 && col != nullptr; }
  // multi-GPU overlap (spmv.cu): bnd_flag[s] = 1 / bnd_list = the slices with a row that references a ghost column; built on first use
  uint8_t *bnd_flag = nullptr;
  int32_t *bnd_list = nullptr;
  // stencil rows + exception rows (stx.cu): bit l of xmask[s] = row 32 s + l is exactly the dominant stencil (distances, values, no ghost column, nothing to
  // push); xrows = all other rows, ascending.  Built on first use; x_comm: built with the multi-GPU exceptions (ghost columns, rows to push)
  uint32_t *xmask = nullptr;
  int32_t *xrows = nullptr;
  int nx = -1, x_comm = 0;
  // ... and the exception rows' entries once more, packed: SELL-32 over the LIST (entry j of list position i at xs_ptr[i >> 5] + 32 j + (i & 31),
  // explicit columns, explicit values) -- gathering them from the matrix' own slices costs a 32-byte sector per 4- or 8-byte entry
  int64_t *xs_ptr = nullptr; uint16_t *xs_len = nullptr; int32_t *xs_col = nullptr; double *xs_val = nullptr;
  int64_t xs_entries = 0;
  int64_t  gen = 0;               // value generation: bumped whenever the values change (sell_update_diag); what was derived from them (base-level LU) checks it
  struct TrcData *trc = nullptr;  // transfer stencils: rows by class (trc.cu), built on first use
  uint8_t *comm_flag = nullptr;   // [slices] HaloK::flag of launches over this matrix' rows (comm.cu halo_comm_flag): bit 0 ghost columns, bit 1 rows to push
  int n_int = -1, n_bnd = 0;
  struct TriSched *tri[2] = {nullptr, nullptr};   // Gauss-Seidel schedules (gs.cu): lower / upper triangle in dependency-level order, built on demand
  int64_t  col_words = 0;         // column words a pass over the matrix fetches from HBM: true entries of explicit slices + the distinct distance tables
  // compulsory bytes of one pass over the stored matrix (values + column words), the "z*W" term of SURVEY.md 8(d) for this format
  double entry_bytes() const { return (vcode ? 1.0 : 8.0) * (double)(val_entries >= 0 ? val_entries : nnz) * bb + 4.0 * (double)col_words; }
};

struct SellView {          // what a kernel needs (passed by value)
  int n;
  int fixed_w;
  const int64_t *slice_ptr;
  const int64_t *col_ptr;
  const uint16_t *rowlen;
  const int32_t *col;
  const double *val;
  const double *diag;
  const uint8_t *vcode;      // transfer stencils with a value table (else nullptr)
  const double *vtable;
  const double *vt;          // shared value tables of uniform slices (else nullptr)
};
static inline SellView view(const SellMat &m) { return SellView{m.n, m.fixed_w, m.slice_ptr, m.col_ptr, m.rowlen, m.col, m.val, m.diag, m.vcode, m.vtable, m.vt}; }

#ifdef __CUDACC__
// entry offset of slice s (and its width): computed for fixed-width matrices -- one dependent load less per row
__device__ __forceinline__ int64_t slice_off(const SellView &A, int s) { return A.fixed_w ? (int64_t)s * 32 * A.fixed_w : A.slice_ptr[s]; }
__device__ __forceinline__ int slice_width(const SellView &A, int s, int64_t sp) { return A.fixed_w ? A.fixed_w : (int)((A.slice_ptr[s + 1] - sp) >> 5); }
// column index of entry j of row r:  __ldg(ci.p + j * ci.stride) + ci.base   (warp-uniform stride/base)
struct ColIter { const int32_t *p; int stride; int base; };
__device__ __forceinline__ ColIter col_iter(const SellView &A, int r)
{
  const int64_t cp = (A.fixed_w && A.col_ptr == A.slice_ptr) ? (int64_t)(r >> 5) * 32 * A.fixed_w : A.col_ptr[r >> 5];
  if (cp < 0) return ColIter{A.col + UG_COLTAB(cp), 1, r};
  return ColIter{A.col + cp + (r & 31), 32, 0};
}
__device__ __forceinline__ int col_at(const ColIter &ci, int j) { return __ldg(ci.p + (size_t)j * ci.stride) + ci.base; }
#endif

// ---- multi-GPU halo, peer-memory ghost rows (comm.cu) ----------------------------------------------------------------
// Every kernel that reads ghost columns or produces a vector whose interface rows the neighbours need is a "comm kernel".  All ranks
// launch the same comm kernels in the same order and number them g = 1, 2, ... (Comm::kseq).  At its head block 0 publishes g into the
// `started` word it owns in every neighbour's flag block; the warps whose slice reads ghost columns or holds rows to push wait until
// every neighbour's `started` word has reached g -- then everything the neighbours launched before kernel g is complete: their pushes
// have landed in my ghost rows, and their reads of the ghost rows I am about to overwrite are done.  Rows to push are stored straight
// into the ghost rows of the neighbours' copy of the same vector (mapped with CUDA IPC), by the kernel that computes them.
#define HALO_MAX_NB 26
struct HaloDev {                                         // device-resident tables of one partitioned level
  int nnb;
  const unsigned long long *my_flag[HALO_MAX_NB];        // neighbour k's words in MY flag block: [0] started, [1] landed
  unsigned long long *peer_flag[HALO_MAX_NB];            // my words in neighbour k's flag block (peer memory)
  const uint32_t *snd_bits;                              // [slices] lanes whose row goes to at least one neighbour
  const int32_t *snd_first;                              // [slices] number of send rows in the slices before this one
  const int32_t *srow_ptr;                               // [send rows + 1] -> snd_ent
  const uint32_t *snd_ent;                               // (neighbour index << 27) | row index in that neighbour's ghost region
};
struct HaloK {                                           // kernel parameter; flag == nullptr: not a comm launch
  const uint8_t *flag;                                   // per slice of the kernel's rows: bit 0 reads ghost columns, bit 1 holds rows to push
  const HaloDev *cdev, *pdev;                            // level whose ghost rows are read / level whose rows are pushed (either may be nullptr)
  unsigned long long g;                                  // number of this comm kernel
  double *const *peer;                                   // [pdev->nnb] the neighbours' ghost regions of the vector this kernel produces; nullptr: no push
  int *err;
  int sel;                                               // smoothing step: which of its results is pushed (HALO_PUSH_*)
  unsigned long long *go;                                // [HALO_GO_SLOTS * 16] local "all neighbours have started kernel g" words, one 128-byte line per SM
};
#define HALO_GO_SLOTS 256
#define HALO_DBG_NOPUSH 256     // timing experiments only (UGGPU_DBG_HALO): bits of HaloK::sel that switch the stores / the waits off
#define HALO_DBG_NOWAIT 512
enum { HALO_PUSH_NONE = 0, HALO_PUSH_TOUT = 1, HALO_PUSH_B = 2, HALO_PUSH_C = 3 };
static inline HaloK halo_none() { return HaloK{nullptr, nullptr, nullptr, 0ull, nullptr, nullptr, 0, nullptr}; }
// what the caller of a comm-aware kernel knows about the exchange around it (cycle.cu: the fused schedule)
struct HaloPlan {
  bool operand_ready;      // the ghost rows of the kernel's operand were pushed by the kernel that produced it: no exchange before this launch
  double *push;            // vector this launch produces whose interface rows the next kernel's ghost columns need (nullptr: none)
  int push_level;          // its level
};

#ifdef __CUDACC__
__device__ __forceinline__ void halo_publish_dev(const HaloDev *d, unsigned long long g, int word)
{
  if ((int)threadIdx.x < d->nnb) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(d->peer_flag[threadIdx.x] + word), "l"(g) : "memory");
  }
}
__device__ __forceinline__ void halo_wait_dev(const HaloDev *d, unsigned long long g, int word, int *err);
// Block 0 of a comm kernel: "everything this rank launched before kernel g is complete" goes to the neighbours; then its first warp
// is the WATCHER: it waits for the neighbours' words and raises the local go words, one per SM.  The comm warps poll only their SM's
// go word -- hundreds of thousands of warps polling the neighbours' words themselves serialise on one L2 address (measured: +1 ms per
// launch on a 4*10^8-row level).
__device__ __forceinline__ void halo_publish(const HaloK &h)
{
  if (blockIdx.x != 0) return;
  if (h.cdev) halo_publish_dev(h.cdev, h.g, 0);
  if (h.pdev && h.pdev != h.cdev) halo_publish_dev(h.pdev, h.g, 0);
  if (threadIdx.x < 32) {
    if (h.cdev) halo_wait_dev(h.cdev, h.g, 0, h.err);
    if (h.pdev && h.pdev != h.cdev) halo_wait_dev(h.pdev, h.g, 0, h.err);
    for (int s = threadIdx.x; s < HALO_GO_SLOTS; s += 32) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(h.go + (size_t)s * 16), "l"(h.g) : "memory");
  }
}
__device__ __forceinline__ void halo_wait_dev(const HaloDev *d, unsigned long long g, int word, int *err)      // one warp
{
  const int lane = threadIdx.x & 31;
  if (lane < d->nnb) {
    unsigned long long t0 = 0, t1, got;
    for (int it = 0;; it++) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(d->my_flag[lane] + word) : "memory");
      if (got >= g) break;
      if (it == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if ((it & 63) == 63) {
        if (*reinterpret_cast<volatile int *>(err)) break;                 // an earlier wait already failed: do not wait again
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 10000000000ull) { atomicExch(err, UGGPU_CUDA_ERROR); break; }   // 10 s: report instead of hanging the GPU
      }
      __nanosleep(100);
    }
  }
  __syncwarp();
}
// a warp whose slice reads ghost columns or holds rows to push (whole warp): wait for this SM's go word
__device__ __forceinline__ void halo_wait(const HaloK &h)
{
  if (h.sel & HALO_DBG_NOWAIT) return;
  if ((threadIdx.x & 31) == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned long long *w = h.go + (size_t)(smid & (HALO_GO_SLOTS - 1)) * 16;
    unsigned long long t0 = 0, t1, got;
    for (int it = 0;; it++) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(got) : "l"(w) : "memory");
      if (got >= h.g) break;
      if (it == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      if ((it & 63) == 63) {
        if (*reinterpret_cast<volatile int *>(h.err)) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 12000000000ull) { atomicExch(h.err, UGGPU_CUDA_ERROR); break; }
      }
      __nanosleep(200);
    }
  }
  __syncwarp();
}
// row r of the produced vector -> the ghost rows of the neighbours that hold a copy of it
template <int BS>
__device__ __forceinline__ void halo_push_row(const HaloK &h, int r, const double (&v)[BS])
{
  if (h.sel & HALO_DBG_NOPUSH) return;
  const HaloDev *d = h.pdev;
  const uint32_t bits = d->snd_bits[r >> 5];
  const int lane = r & 31;
  if (!((bits >> lane) & 1u)) return;
  const int i = d->snd_first[r >> 5] + __popc(bits & ((1u << lane) - 1u));
  for (int e = d->srow_ptr[i]; e < d->srow_ptr[i + 1]; e++) {
    const uint32_t ent = d->snd_ent[e];
    double *dst = h.peer[ent >> 27] + (size_t)(ent & 0x7ffffffu) * BS;
#pragma unroll
    for (int q = 0; q < BS; q++) dst[q] = v[q];
  }
}
#endif

struct Level {
  bool exists = false;
  int n = 0, bs = 0;
  uint8_t *vclass = nullptr, *vnclass = nullptr, *ctl = nullptr;
  uint32_t *skip = nullptr;
  std::map<int, SellMat> mats;
  std::map<int, double *> vecs;
  std::map<int, cudaEvent_t> pending;   // vectors with an asynchronous upload in flight on the copy stream (uggpu_vec_upload_async)
  SellMat P, R;                  // P: rows = this level, cols = level-1;  R: rows = level-1, cols = this level
  int transfer_mode = 0;         // UGGPU_TRANSFER_STANDARD / UGGPU_TRANSFER_IMAT (damping applied after the sums, transfer.cu)
  // base-level dense LU (column-major, inverse diagonal stored), built by uggpu_lmgc_preprocess
  double *lu = nullptr;
  int luN = 0, luA = -1;
  int64_t luGen = -1;            // value generation of the matrix the factorisation was made from
  // scalar levels: the rows' lower / upper entries in the order of UG's matrix lists after l_lrdecomp's fill-in (cycle.cu lu_lists)
  int32_t *lu_lo_ptr = nullptr, *lu_lo_col = nullptr, *lu_up_ptr = nullptr, *lu_up_col = nullptr, *lu_lo_row = nullptr, *lu_up_row = nullptr;
  double *lu_lo_val = nullptr, *lu_up_val = nullptr, *lu_dinv = nullptr;
  int lu_lo_nnz = 0, lu_up_nnz = 0, lu_active = 0;
  // multi-GPU (part.h, comm.cu): n = rows this rank owns; vectors carry nghost extra rows at the tail
  int nghost = 0;
  bool partitioned = false;      // rows are split over the ranks (halo exchange + global reductions apply)
  int64_t n_global = 0;          // rows of the whole level over all ranks
  struct PartGrid *part = nullptr;      // host copy
  struct PartGrid *d_part = nullptr;    // device copy
  int32_t *d_send_idx = nullptr;        // owned rows to pack, grouped by neighbour (PartGrid::nb_send_off)
  int send_total = 0;
  std::vector<int> peer_recv_off;       // per neighbour: where my rows start in ITS ghost region (peer-memory halo exchange, comm.cu)
  // the level's interface lists in the form every transport uses (filled by the synthetic generator from PartGrid, or by
  // uggpu_level_set_partition from the caller's lists): neighbour ranks, send list offsets into d_send_idx, receive offsets into my ghost rows
  int nnb = 0;
  std::vector<int> nb_rank, nb_send_off, nb_recv_off;
  struct LevelHalo *halo = nullptr;     // comm.cu: peer-memory ghost rows (tables, mapped neighbour vectors)
  double *last_pushed = nullptr;        // vector whose interface rows the last producing kernel stored into the neighbours' ghost rows
};

struct uggpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;   // second stream for uploads that overlap the cycle (created on first use)
  cudaStream_t halo_stream = nullptr;   // interface rows of a partitioned level: wait for the halo + compute, next to the interior rows (created on first use)
  cudaEvent_t halo_ev[2] = {nullptr, nullptr};
  Level lev[UGGPU_MAX_LEVELS];
  int fullrefinelevel = 0;
  int64_t launches = 0;
  int64_t bytes = 0;
  int64_t value_gen = 0;       // source of SellMat::gen (unique per context, so a re-created matrix never repeats a number)
  // reduction scratch
  double *partials = nullptr;  size_t partials_cap = 0;   // device
  double *dres = nullptr;                                  // device results [UGGPU_MAX_LEVELS*4*UGGPU_MAX_BS]
  double *hres = nullptr;                                  // pinned host mirror
  int *derr = nullptr;                                     // device error word
  int *herr = nullptr;                                     // pinned
  int sm_count = 148;
  // per-kernel event timing (uggpu_prof_*): records are resolved lazily after a synchronise
  bool prof = false;
  struct ProfRec { int kind, level; double bytes; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  // multi-GPU (comm.cu)
  void *comm = nullptr;
};

// RAII bracket around one kernel launch: two events on the context's stream when profiling is on
struct ProfScope {
  uggpu_ctx *ctx; bool on;
  uggpu_ctx::ProfRec rec;
  ProfScope(uggpu_ctx *c, int kind, int level, double bytes);
  ~ProfScope();
};

// ---- software prefetch into L2 ---------------------------------------------------------------------------------------
// A thread-per-row warp lives for ~6 dependent memory round trips (slice offsets, then the row's entries in groups of
// four); with every one of them an HBM miss the SMs run out of warps long before HBM runs out of bandwidth (ncu on the
// smoothing kernel: long-scoreboard stalls, 59 % DRAM throughput).  Every warp therefore ends by touching, with one
// prefetch.global.L2 per 128-byte line, the value block (and explicit column words, and vector entries) of the slice
// `dist` slices ahead -- one generation of resident warps -- so that that generation finds its streams in L2.
struct Prefetch {
  int dist;        // slices ahead (0: off)
  int nsl;         // slices of the matrix
  int mode;        // bit 0 values, bit 1 explicit column words, bit 2 vector entries of the rows, bit 3 slice offsets / row lengths two hops ahead, bit 4 row flags,
                   // bit 5 transfer kernels on fixed-width stencils touch the far lines when the warp starts (else when it ends)
  int val_lines;   // 128-byte lines of the widest slice's value block
  int col_lines;   // same for explicit column words
  int64_t val_bytes, col_bytes, vec_bytes;   // sizes of the arrays: no line beyond them is touched
};
// ctx.cu; bs: components per row of the vectors; slices_per_warp: slices one warp of the kernel works on; UGGPU_PF_DIST / UGGPU_PF_MODE override
Prefetch make_prefetch(const uggpu_ctx *ctx, const SellMat *A, int bs, int slices_per_warp = 1);

#ifdef __CUDACC__
struct PfState { int64_t sp, cp; int slice; };
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// offsets of the far slice: requested at the start of the row's work ...
__device__ __forceinline__ PfState pf_begin(const SellView &A, int r, const Prefetch &pf)
{
  PfState st{-1, -1, (r >> 5) + pf.dist};
  // two hops: the slice offsets / row lengths of the slice 2*dist ahead are direct-indexed -> touched now, so that the loads of
  // the far slice's offsets below (and every kernel's first loads) are L2 hits, not HBM round trips
  const int far2 = (r >> 5) + 2 * pf.dist, lane = threadIdx.x & 31;
  if ((pf.mode & 8) && pf.dist > 0 && far2 < pf.nsl && lane < 3) {
    if (lane == 2) prefetch_l2(A.rowlen + (size_t)far2 * 32);
    else if (!A.fixed_w || (lane == 1 && A.col_ptr != A.slice_ptr)) prefetch_l2(lane == 0 ? A.slice_ptr + far2 : A.col_ptr + far2);
  }
  if (pf.dist > 0 && st.slice < pf.nsl) {
    st.sp = A.fixed_w ? (int64_t)st.slice * 32 * A.fixed_w : __ldg(A.slice_ptr + st.slice);
    if ((pf.mode & 2) || A.vt) st.cp = (A.fixed_w && A.col_ptr == A.slice_ptr) ? st.sp : __ldg(A.col_ptr + st.slice);
  }
  return st;
}
// ... and used at its end: lane l touches lines l, l + 32, ... of the far slice's value block and line l of its column words
template <int BB>
__device__ __forceinline__ void pf_end(const SellView &A, const PfState &st, const Prefetch &pf)
{
  if (st.sp < 0) return;
  const int lane = threadIdx.x & 31;
  if ((pf.mode & 1) && !(A.vt && st.cp < 0 && UG_VALTAB(st.cp) >= 0)) {      // a slice with shared values has no value stream
    const int64_t o = st.sp * BB * (int64_t)sizeof(double);
    for (int l = lane; l < pf.val_lines; l += 32)
      if (o + (int64_t)l * 128 < pf.val_bytes) prefetch_l2(reinterpret_cast<const char *>(A.val) + o + (int64_t)l * 128);
  }
  if ((pf.mode & 2) && st.cp >= 0 && lane < pf.col_lines) {
    const int64_t o = st.cp * (int64_t)sizeof(int32_t) + (int64_t)lane * 128;
    if (o < pf.col_bytes) prefetch_l2(reinterpret_cast<const char *>(A.col) + o);
  }
}
// the far slice's 32 rows of a vector with BS components per row
template <int BS>
__device__ __forceinline__ void pf_vec(const double *v, const PfState &st, const Prefetch &pf)
{
  const int lane = threadIdx.x & 31;
  if (lane < 2 * BS) {
    const int64_t o = ((int64_t)st.slice * 32 * BS) * (int64_t)sizeof(double) + (int64_t)lane * 128;
    if (o < pf.vec_bytes) prefetch_l2(reinterpret_cast<const char *>(v) + o);
  }
}
// the far slice's 32 rows of a per-row array of ELEM-byte entries (flags)
template <int ELEM>
__device__ __forceinline__ void pf_rows(const void *base, const PfState &st, const Prefetch &pf)
{
  const int lane = threadIdx.x & 31;
  if ((pf.mode & 16) && lane < (32 * ELEM + 127) / 128 && st.slice < pf.nsl) prefetch_l2(reinterpret_cast<const char *>(base) + ((size_t)st.slice * 32 + 0) * ELEM + (size_t)lane * 128);
}
#endif

#ifdef __CUDACC__
// ---- small dense blocks (np/algebra/block.cc) ----------------------------------------------------------------------
// InvertSmallBlock, np/algebra/block.cc:272-321 (closed forms for n = 2, 3); returns non-zero for det == 0
template <int BS>
__device__ __forceinline__ int invert_small_block(const double *mat, double *inv)
{
  if (BS == 2) {
    double det = mat[0] * mat[3] - mat[1] * mat[2];
    if (det == 0.0) return 1;
    double invdet = 1.0 / det;
    inv[0] = mat[3] * invdet; inv[1] = -mat[1] * invdet; inv[2] = -mat[2] * invdet; inv[3] = mat[0] * invdet;
    return 0;
  }
  double det = mat[0] * mat[4] * mat[8 % (BS * BS)] + mat[1] * mat[5 % (BS * BS)] * mat[6 % (BS * BS)] + mat[2] * mat[3] * mat[7 % (BS * BS)]
               - mat[2] * mat[4] * mat[6 % (BS * BS)] - mat[0] * mat[5 % (BS * BS)] * mat[7 % (BS * BS)] - mat[1] * mat[3] * mat[8 % (BS * BS)];
  if (det == 0.0) return 1;
  double invdet = 1.0 / det;
  constexpr int BB = BS * BS;
  inv[0] = ( mat[4 % BB] * mat[8 % BB] - mat[5 % BB] * mat[7 % BB]) * invdet;
  inv[3] = (-mat[3] * mat[8 % BB] + mat[5 % BB] * mat[6 % BB]) * invdet;
  inv[6 % BB] = ( mat[3] * mat[7 % BB] - mat[4 % BB] * mat[6 % BB]) * invdet;
  inv[1] = (-mat[1] * mat[8 % BB] + mat[2] * mat[7 % BB]) * invdet;
  inv[4 % BB] = ( mat[0] * mat[8 % BB] - mat[2] * mat[6 % BB]) * invdet;
  inv[7 % BB] = (-mat[0] * mat[7 % BB] + mat[1] * mat[6 % BB]) * invdet;
  inv[2] = ( mat[1] * mat[5 % BB] - mat[2] * mat[4 % BB]) * invdet;
  inv[5 % BB] = (-mat[0] * mat[5 % BB] + mat[2] * mat[3]) * invdet;
  inv[8 % BB] = ( mat[0] * mat[4 % BB] - mat[1] * mat[3]) * invdet;
  return 0;
}

// SolveSmallBlock (block.cc:104-142), n = 1,2,3.  Returns non-zero for a singular 2x2 block.
template <int BS>
__device__ __forceinline__ int solve_small_block(const double (&mat)[BS * BS], const double (&rhs)[BS], double (&sol)[BS])
{
  if (BS == 1) { sol[0] = rhs[0] / mat[0]; return 0; }
  if (BS == 2) {
    double det = mat[0] * mat[3 % (BS * BS)] - mat[1 % (BS * BS)] * mat[2 % (BS * BS)];
    if (det == 0.0) return 1;
    det = 1.0 / det;
    sol[0] = (rhs[0] * mat[3 % (BS * BS)] - rhs[1 % BS] * mat[1 % (BS * BS)]) * det;
    sol[1 % BS] = (rhs[1 % BS] * mat[0] - rhs[0] * mat[2 % (BS * BS)]) * det;
    return 0;
  }
  // n == 3 (indices wrapped with % only to keep the BS<3 instantiations well-formed)
  constexpr int BB = BS * BS;
  double M3div0 = mat[3 % BB] / mat[0];
  double M6div0 = mat[6 % BB] / mat[0];
  double aux = (mat[7 % BB] - M6div0 * mat[1 % BB]) / (mat[4 % BB] - M3div0 * mat[1 % BB]);
  sol[2 % BS] = (rhs[2 % BS] - M6div0 * rhs[0] - aux * (rhs[1 % BS] - M3div0 * rhs[0]))
                / (mat[8 % BB] - M6div0 * mat[2 % BB] - aux * (mat[5 % BB] - M3div0 * mat[2 % BB]));
  sol[1 % BS] = (rhs[1 % BS] - mat[3 % BB] / mat[0] * rhs[0] - (mat[5 % BB] - M3div0 * mat[2 % BB]) * sol[2 % BS])
                / (mat[4 % BB] - M3div0 * mat[1 % BB]);
  sol[0] = (rhs[0] - mat[1 % BB] * sol[1 % BS] - mat[2 % BB] * sol[2 % BS]) / mat[0];
  return 0;
}

// C = A * B for bs x bs blocks with the reference's summation (sum = 0; sum += a*b, ugiter.cc:3814-3822); returns true if C == 0
template <int BS>
__device__ __host__ __forceinline__ bool block_mul(const double *a, const double *b, double *c)
{
  bool zero = true;
  for (int i0 = 0; i0 < BS; i0++)
    for (int j0 = 0; j0 < BS; j0++) {
      double sum = 0.0;
      for (int k0 = 0; k0 < BS; k0++) sum += a[i0 * BS + k0] * b[k0 * BS + j0];
      c[i0 * BS + j0] = sum;
      if (sum != 0.0) zero = false;
    }
  return zero;
}

#endif

// ---- error plumbing -------------------------------------------------------------------------------
int uggpu_fail(int code, const char *fmt, ...);
#define CUDA_TRY(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t e__ = (expr);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return uggpu_fail(UGGPU_CUDA_ERROR, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
  } while (0)
#define UG_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)
#define KCHECK(ctx) do { (ctx)->launches++; CUDA_TRY(cudaGetLastError()); } while (0)

// ---- helpers (ctx.cu) -------------------------------------------------------------------------------
int dev_alloc(uggpu_ctx *ctx, void **p, size_t bytes);
int dev_free(uggpu_ctx *ctx, void *p, size_t bytes);
template <class T> static inline int dalloc(uggpu_ctx *ctx, T **p, size_t count) { return dev_alloc(ctx, (void **)p, count * sizeof(T)); }
template <class T> static inline int dfree(uggpu_ctx *ctx, T *&p, size_t count) { int rc = dev_free(ctx, (void *)p, count * sizeof(T)); p = nullptr; return rc; }
Level *get_level(uggpu_ctx *ctx, int level);                 // NULL + error if absent
double *get_vec(uggpu_ctx *ctx, int level, int vec);         // NULL + error if absent; orders the compute stream behind a pending upload of the vector
double *get_vec_lazy(uggpu_ctx *ctx, int level, int vec);    // the pointer only: the caller calls vec_wait() before the first kernel that touches it
int vec_wait(uggpu_ctx *ctx, int level, int vec);
SellMat *get_mat(uggpu_ctx *ctx, int level, int mat);
SellMat *get_mat_quiet(uggpu_ctx *ctx, int level, int mat);
int ensure_partials(uggpu_ctx *ctx, size_t count);
int check_device_error(uggpu_ctx *ctx);                      // sync + read the device error word

// ---- SELL (sell.cu) -----------------------------------------------------------------------------------
// Slice offsets from slice widths.  When padding every slice to the widest one costs at most 1/8 more storage the matrix is
// laid out with that one width (*fixed_w > 0; UGGPU_NO_FIXED_WIDTH=1 disables): SELL degenerates to ELL and kernels compute
// the offsets.  Padding is never read, so results do not depend on the choice.
void sell_layout(const std::vector<int> &width, std::vector<int64_t> &sp, int *maxlen, int *fixed_w);
// Builds a SELL-32 matrix from device CSR (rowptr int64[n+1], col, val[nnz*bb] row-major blocks).
int sell_from_device_csr(uggpu_ctx *ctx, int n, int bb, const int64_t *d_rowptr, const int32_t *d_col, const double *d_val, SellMat *out);
int sell_from_host_csr(uggpu_ctx *ctx, int n, int bb, const int32_t *rowptr, const int32_t *col, const double *val, SellMat *out);
int sell_set_values_host(uggpu_ctx *ctx, SellMat *m, const double *val);
int sell_to_host_csr(uggpu_ctx *ctx, const SellMat *m, int32_t *rowptr, int32_t *col, double *val);
int sell_free(uggpu_ctx *ctx, SellMat *m);
int sell_clone(uggpu_ctx *ctx, const SellMat *src, SellMat *dst);    // deep copy: pattern, layout, values (AllocMDFromMD + dmatcopy)
int sell_free_schedules(uggpu_ctx *ctx, SellMat *m);          // gs.cu: drops the Gauss-Seidel schedules (they hold a copy of the values)
// replaces the explicit column words of uniform slices by one distance per slice column (lossless); no-op when nothing is gained
int sell_compress_cols(uggpu_ctx *ctx, SellMat *m);
// (re)builds m->diag from the values (after the matrix is built and after every change of its values)
int sell_update_diag(uggpu_ctx *ctx, SellMat *m);
// scalar-entry matrices with at most 256 distinct stored values get a one-byte code per entry + a table (lossless); no-op otherwise
int sell_compress_values(uggpu_ctx *ctx, SellMat *m);
// matrices: uniform slices whose rows also hold identical values share value tables (lossless; the explicit values stay); drop = forget them
int sell_share_values(uggpu_ctx *ctx, SellMat *m);
int sell_drop_shared_values(uggpu_ctx *ctx, SellMat *m);

// ---- kernels used across files --------------------------------------------------------------------------
struct Damp { double a[UGGPU_MAX_BS]; };
static inline Damp mkdamp(const double *d, int bs) { Damp r; for (int i = 0; i < UGGPU_MAX_BS; i++) r.a[i] = (d && i < bs) ? d[i] : 1.0; return r; }

// rowmode: 0 all rows, 1 NEW_DEFECT rows, 2 FINE_GRID_DOF rows (vecloop.ct:22-49)
int k_dmatmul(uggpu_ctx *ctx, int level, int op, int rowmode, int x, int M, int y);
int k_vec_op(uggpu_ctx *ctx, int level, int rowmode, int op, double *x, const double *y, Damp a);
int k_reduce(uggpu_ctx *ctx, int level, int rowmode, int kind, const double *x, const double *y, int slot);
int fetch_results(uggpu_ctx *ctx, int nslots);   // dres -> hres, synchronises
int reduce_partials_final(uggpu_ctx *ctx, int bs, size_t count, int slot, int level);   // ctx->partials -> dres[slot] (+ all-reduce on partitioned levels)
struct LoopItem { int level, rowmode; };
// (level, rowmode) pairs of one reference loop over levels fl..tl in `mode` (vecloop.ct:22-49)
int surface_loop(uggpu_ctx *ctx, int fl, int tl, int mode, std::vector<LoopItem> &out);
int reduce_loop(uggpu_ctx *ctx, int fl, int tl, int mode, int kind, int x, int y, double *sums, int *bs_out);

enum { VOP_SET, VOP_COPY, VOP_SCALX, VOP_ADD, VOP_SUB, VOP_MINUSADD, VOP_AXPYX };
enum { RED_DOT, RED_NRM2 };
// Fused BLAS-1 chains of the Krylov solvers (blas1.cu chain_loop): several reference calls on the same rows in ONE pass, every entry
// receiving the same operations in the same order (results bit-identical to the separate calls), optionally with the partial sums of
// the reduction that follows (same launch geometry and accumulation order as k_red_rows, so the sums are bit-identical, too).
enum { CH_ADD_DOT,      // v0 += v1;                                   sum v2 * v0                      (cg: dadd b t, ddot c b)
       CH_SCAL_ADD,     // v0 = v0 * a0 + v1                                                            (cg: dscal p, dadd p c)
       CH_AXPY2_NRM,    // v0 += a0 * v1;  v2 += a1 * v3;               sum v2 * v2                      (cg: daxpy x p, daxpy b t, residuum)
       CH_BCGS_P,       // v0 = v0 * a0 + v1 + a1 * v2;  v3 = 0;  v4 = v0                                (bcgs: dscal p, dadd p b, daxpy p v, dset q, dcopy s p)
       CH_BCGS_S,       // v0 += a0 * v1;  v2 = v3 + a1 * v4;           sum v2 * v2                      (bcgs: daxpy x q, dcopy s b, daxpy s v, residuum of s)
       CH_COUNT };
// levels fl..tl, every row (ALL_VECTORS); the reduction, if the chain has one, over the ON_SURFACE rows of those levels; sums[bs] summed
// over the levels in level order on the host like reduce_loop
int chain_loop(uggpu_ctx *ctx, int fl, int tl, int chain, const int *vecs, double a0, double a1, double *sums, int *bs_out);

// smoother step flags (spmv.cu, DESIGN.md "fused kernels")
enum { SF_TOUT = 1, SF_CADD = 2, SF_CSET = 4, SF_XADD = 8, SF_NORM = 16,
       // The step BEFORE this one left c alone (no SF_CADD / SF_CSET); this one adds both corrections, in the reference's order:
       // c = (c + t_prev) + t_in.  t_prev is the previous step's input, which sits in the row of the buffer this step writes its output to
       // (the two temporaries alternate) and is read by the row's own thread before it is overwritten.  Saves one read + write of c per
       // pair of steps (16 B per row; 8 B come back as the read of t_prev).  Only the stencil-rows / exception-rows kernels (stx.cu)
       // take it: cycle.cu asks stx_handles() before it schedules a pair.
       SF_CPREV = 32 };
// hp (may be nullptr): what the fused schedule knows about the exchanges around the launch (HaloPlan); without it every launch
// exchanges its operand's ghost rows itself and pushes nothing
int k_smooth_step(uggpu_ctx *ctx, int level, int A, int flags, const double *tin, double *b, double *c, double *tout,
                  Damp damp, double *x, int norm_slot, const HaloPlan *hp = nullptr);
int k_jac(uggpu_ctx *ctx, int level, int A, double *v, const double *d, Damp damp, const HaloPlan *hp = nullptr);
// stx.cu: the smoothing step / dmatmul of a matrix with a dominant stencil as TWO kernels -- all rows that are exactly the stencil, and the
// compact list of all other rows (boundary rows, rows with ghost columns, rows to push).  *done = 0: not applicable, the caller takes its own kernels.
int stx_smooth(uggpu_ctx *ctx, Level *L, SellMat *A, int flags, const double *tin, double *b, double *c, double *tout, Damp damp, double *x, int norm_slot,
               const HaloK &hk, int *done);
int stx_dmatmul(uggpu_ctx *ctx, Level *L, SellMat *A, int op, uint8_t bit, double *x, const double *y, int *done);
int stx_free(uggpu_ctx *ctx, SellMat *m);
bool stx_handles(const Level *L, const SellMat *A);            // smoothing steps of this matrix go to the stx kernel pair (any flag combination of stx_smooth1)
double stx_matrix_bytes(const Level *L, const SellMat *A);   // matrix bytes one pass of the stx kernel pair fetches (< 0: not applicable / not built)
// trc.cu: restriction / interpolation on stencils whose rows fall into a few classes (base column + class byte per row), exception rows as a
// second kernel.  *done = 0: not applicable, the caller launches transfer.cu's kernels.
int trc_interpolate(uggpu_ctx *ctx, Level *F, Level *C, double *to, const double *from, Damp damp, const HaloK &hk, int *done);
int trc_restrict(uggpu_ctx *ctx, Level *F, Level *C, double *to, const double *from, Damp damp, bool fuse, const SellMat *Ac, double *tout, double *czero, Damp sdamp,
                 const HaloK &hk, int *done);
int trc_free(uggpu_ctx *ctx, SellMat *m);
int trc_free_comm(uggpu_ctx *ctx, SellMat *m);       // only if it was built with the multi-GPU exceptions
const uint32_t *halo_snd_bits(const Level *L);   // comm.cu: per slice, the lanes whose row is pushed to a neighbour (nullptr: no tables)   // v = damp * Diag(A)^-1 d (class-masked)
// transfer.cu: fine `level` -> level-1; with fuse: also tout = sdamp*Diag(A_{level-1})^-1 to, czero = 0 on level-1
int k_restrict(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp, bool fuse, int A, double *tout, double *czero, Damp sdamp,
               const HaloPlan *hp = nullptr);
int k_interpolate(uggpu_ctx *ctx, int level, double *to, const double *from, Damp damp, const HaloPlan *hp = nullptr);
// comm.cu: all are no-ops (return 0) when the context has no communicator or the level is not partitioned
// Prepares a comm-aware launch over the rows of matrix M (rows on `row_level`; its columns index vectors of `col_level`, whose ghost rows
// the launch reads -- -1: none): exchanges the operand's ghost rows first unless the plan says its producer pushed them, registers the
// vector to push, numbers the launch.  *hk stays halo_none() when the peer-memory ghost transport does not apply (one GPU, levels held
// completely, IPC unavailable: then the operand has been exchanged through the window / NCCL path and nothing is pushed).
int halo_prepare(uggpu_ctx *ctx, int row_level, int col_level, SellMat *M, double *operand, const HaloPlan *hp, HaloK *hk);
bool halo_fused_available(uggpu_ctx *ctx, int level);                    // the peer-memory ghost transport is (or can be) used on this level
int halo_exchange(uggpu_ctx *ctx, int level, double *v);                 // owned values -> the neighbours' ghost rows of v
int halo_begin(uggpu_ctx *ctx, int level, double *v, int *split);        // push only; *split = 1: finish with halo_finish on another stream
int halo_finish(uggpu_ctx *ctx, int level, double *v, cudaStream_t st);  // wait + unpack on st
int allreduce_sum(uggpu_ctx *ctx, double *dptr, size_t count);           // in place, on the context's stream
int level_free_part(uggpu_ctx *ctx, Level *L);
bool halo_vec_release(uggpu_ctx *ctx, Level *L, double *p, size_t bytes);   // true: other ranks have the vector mapped, it is parked, do not free it
int level_free_lu(uggpu_ctx *ctx, Level *L);                              // cycle.cu: the base-level factorisation
static inline size_t vec_count(const Level *L) { return ((size_t)L->n + (size_t)L->nghost) * (size_t)L->bs; }
// internal temporary vector handles (never visible through the C-ABI callers' handle space)
#define UGGPU_VEC_TMP_A (-1001)
#define UGGPU_VEC_TMP_B (-1002)
#define UGGPU_VEC_TMP_C (-1003)

#endif
