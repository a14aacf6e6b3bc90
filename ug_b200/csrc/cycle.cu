// cycle.cu -- the multigrid cycle Lmgc (np/procs/iter.cc:7741-7949), its base-level solver `ls $I lu`
// (np/procs/ls.cc:637-749 around l_lrdecomp/l_luiter, np/algebra/ugiter.cc:3657,4444) and the outer linear
// solver LinearDefect/LinearResiduum/LinearSolver (np/procs/ls.cc:562-749), all device-resident.
//
// Two schedules produce bit-identical vectors:
//   fused = 0  one kernel per reference call, in the reference's order (the op-for-op mirror);
//   fused = 1  per level and cycle: 1 Jacobi-diagonal pass (fused into the restriction on coarse levels),
//              nu1 + 1 + nu2 fused smoothing-step kernels (spmv.cu k_smooth_k), 1 restriction, 1 prolongation;
//              on the solver's top level the last step also applies x += c and accumulates the defect norm.
#include "uggpu_internal.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

// ---- dense LU of the base level ---------------------------------------------------------------------------------
#define LU_THREADS 1024
#define LU_MAX_N 2048

__global__ void k_lu_scatter(SellView A, int bs, int N, double *__restrict__ lu)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  int bb = bs * bs, lane = r & 31;
  int64_t sp = A.slice_ptr[r >> 5];
  int len = A.rowlen[r];
  const ColIter ci = col_iter(A, r);
  for (int j = 0; j < len; j++) {
    int c = col_at(ci, j);
    for (int i = 0; i < bs; i++)
      for (int k = 0; k < bs; k++)
        lu[(size_t)(c * bs + k) * N + (r * bs + i)] = A.val[(sp + (int64_t)j * 32) * bb + (int64_t)(i * bs + k) * 32 + lane];
  }
}

// Right-looking LU without pivoting in vector-index order; element (i,j) at lu[j*N+i]; the inverse of the
// diagonal is stored (StoreInverse, ugiter.cc:139).  Rows/columns of vectors with VCLASS < ACTIVE_CLASS are left out.
// Every element receives its updates for i = 0,1,2,... in the same order as a sequential elimination.
__global__ void __launch_bounds__(LU_THREADS) k_lu_factor(int N, int bs, const uint8_t *__restrict__ vclass, double *__restrict__ lu)
{
  __shared__ double sinv;
  const int tid = threadIdx.x;
  for (int i = 0; i < N; i++) {
    if (vclass[i / bs] < 3) continue;
    if (tid == 0) { double inv = 1.0 / lu[(size_t)i * N + i]; lu[(size_t)i * N + i] = inv; sinv = inv; }
    __syncthreads();
    const double inv = sinv;
    for (int j = i + 1 + tid; j < N; j += LU_THREADS)
      if (vclass[j / bs] >= 3) lu[(size_t)i * N + j] = lu[(size_t)i * N + j] * inv;
    __syncthreads();
    const int m = N - i - 1;
    for (int idx = tid; idx < m * m; idx += LU_THREADS) {
      int j = i + 1 + idx % m, k = i + 1 + idx / m;
      if (vclass[j / bs] < 3 || vclass[k / bs] < 3) continue;
      double piv = lu[(size_t)i * N + j];
      if (piv == 0.0) continue;
      lu[(size_t)k * N + j] = lu[(size_t)k * N + j] - piv * lu[(size_t)k * N + i];
    }
    __syncthreads();
  }
}

// ---- block rows (bs = 2, 3): the block variant of l_lrdecomp (ugiter.cc:3771-3880) on a dense array of bs x bs blocks -------
// block (row j, column k) at blk[(k*n + j)*bb .. +bb), row-major inside the block (for bs = 1 this is the scalar layout above).
__global__ void k_lu_scatter_block(SellView A, int bs, int n, double *__restrict__ blk)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= A.n) return;
  int bb = bs * bs, lane = r & 31;
  int64_t sp = A.slice_ptr[r >> 5];
  int len = A.rowlen[r];
  const ColIter ci = col_iter(A, r);
  for (int j = 0; j < len; j++) {
    int c = col_at(ci, j);
    for (int q = 0; q < bb; q++) blk[((size_t)c * n + r) * bb + q] = A.val[(sp + (int64_t)j * 32) * bb + (int64_t)q * 32 + lane];
  }
}

// invert_small_block<BS>, block_mul<BS>: uggpu_internal.h (shared with the ILU decomposition in gs.cu)

// Step i: invert the diagonal block and store the inverse (:3784-3793); every block (j,i), j > i, becomes the multiplier
// M_ji * Inv (:3811-3826); every block (j,k), j,k > i, with a non-zero multiplier and a non-zero correction M_ji * M_ik
// loses that correction (:3832-3877).  Blocks outside the pattern are zero, which is what the reference's "create the
// connection when the correction is not zero" amounts to.  Each block receives its corrections for i = 0, 1, 2, ... in order.
template <int BS>
__global__ void __launch_bounds__(LU_THREADS) k_lu_factor_block(int n, const uint8_t *__restrict__ vclass, double *__restrict__ blk, int *err)
{
  constexpr int BB = BS * BS;
  __shared__ double sinv[BB];
  __shared__ uint8_t pz[LU_MAX_N];
  const int tid = threadIdx.x;
  for (int i = 0; i < n; i++) {
    if (vclass[i] < 3) continue;
    if (tid == 0) {
      double *dg = blk + ((size_t)i * n + i) * BB, inv[BB];
      if (invert_small_block<BS>(dg, inv)) { atomicExch(err, UGGPU_SMALL_DIAG); for (int q = 0; q < BB; q++) inv[q] = (q % (BS + 1)) == 0 ? 1.0 : 0.0; }
      for (int q = 0; q < BB; q++) { dg[q] = inv[q]; sinv[q] = inv[q]; }
    }
    __syncthreads();
    for (int j = i + 1 + tid; j < n; j += LU_THREADS) {
      if (vclass[j] < 3) continue;
      double *mji = blk + ((size_t)i * n + j) * BB, piv[BB];
      pz[j] = block_mul<BS>(mji, sinv, piv) ? 1 : 0;
      for (int q = 0; q < BB; q++) mji[q] = piv[q];
    }
    __syncthreads();
    const int m = n - i - 1;
    for (int idx = tid; idx < m * m; idx += LU_THREADS) {
      const int j = i + 1 + idx % m, k = i + 1 + idx / m;
      if (vclass[j] < 3 || vclass[k] < 3 || pz[j]) continue;
      double cor[BB];
      if (block_mul<BS>(blk + ((size_t)i * n + j) * BB, blk + ((size_t)k * n + i) * BB, cor)) continue;
      double *mjk = blk + ((size_t)k * n + j) * BB;
      for (int q = 0; q < BB; q++) mjk[q] = mjk[q] - cor[q];
    }
    __syncthreads();
  }
}

// v = (LU)^-1 d  (l_luiter ugiter.cc:4444): forward sums accumulate column by column (ascending j, the order of a
// sequential row sum), backward sums are formed in ascending j by one thread from products computed in parallel.
__global__ void __launch_bounds__(LU_THREADS) k_lu_solve(int N, int bs, const uint8_t *__restrict__ vclass, const double *__restrict__ lu,
                                                         double *__restrict__ v, const double *__restrict__ d)
{
  __shared__ double vs[LU_MAX_N];
  __shared__ double prod[LU_MAX_N];
  __shared__ uint8_t act[LU_MAX_N];
  const int tid = threadIdx.x;
  constexpr int Q = LU_MAX_N / LU_THREADS;
  double sum[Q];
#pragma unroll
  for (int q = 0; q < Q; q++) sum[q] = 0.0;
  for (int i = tid; i < N; i += LU_THREADS) act[i] = vclass[i / bs] >= 3;
  __syncthreads();
  for (int j = 0; j < N; j++) {
    if ((j % LU_THREADS) == tid) vs[j] = act[j] ? d[j] - sum[j / LU_THREADS] : 0.0;
    __syncthreads();
    if (!act[j]) continue;
    const double vj = vs[j];
#pragma unroll
    for (int q = 0; q < Q; q++) {
      int i = tid + q * LU_THREADS;
      if (i > j && i < N && act[i]) sum[q] += lu[(size_t)j * N + i] * vj;
    }
  }
  __syncthreads();
  for (int i = N - 1; i >= 0; i--) {
    if (!act[i]) continue;
    for (int j = i + 1 + tid; j < N; j += LU_THREADS)
      if (act[j]) prod[j] = lu[(size_t)j * N + i] * vs[j];
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int j = i + 1; j < N; j++) if (act[j]) s += prod[j];
      vs[i] = (vs[i] - s) * lu[(size_t)i * N + i];
    }
    __syncthreads();
  }
  for (int i = tid; i < N; i += LU_THREADS) v[i] = vs[i];
}

// Scalar rows, summed in the order of UG's matrix lists.  l_lrdecomp (ugiter.cc:3657) creates fill-in with
// CreateExtraConnection, which puts the new entries at the SECOND place of both row lists (gm/algebra.cc:1051-1078), and
// l_luiter (ugiter.cc:4470-4518) adds a row's terms in list order -- so the order depends on when each fill-in entry appeared.
// lu_lists() replays that on the host (integers and zero tests only) and packs the factors into two "row programs" (active
// rows ascending with their L entries, active rows descending with their U entries, values gathered by k_lu_pack); the solve
// is ONE WARP: the lanes form a row's products, lane 0 adds them in list order (a sequential sum is the reference's arithmetic),
// and the next row's entries are already in registers while it does.  Forward: v_i = d_i - sum_{c<i} L_ic v_c;  backward:
// v_i = (v_i - sum_{c>i} U_ic v_c) * (1/U_ii).
#define LUS_PRE 4       // entries per lane of the next row held in registers (rows up to 128 entries are fully prefetched)

struct LuProg { const int32_t *row, *ptr, *col; const double *val; int n; };

__global__ void k_lu_pack(int n, int bb, const double *__restrict__ lu, const int32_t *__restrict__ row, const int32_t *__restrict__ ptr, const int32_t *__restrict__ col,
                          double *__restrict__ val, double *__restrict__ dinv)
{
  const int k = blockIdx.x, r = row[k];
  for (int e = ptr[k] * bb + threadIdx.x; e < ptr[k + 1] * bb; e += blockDim.x) val[e] = lu[((size_t)col[e / bb] * n + r) * bb + e % bb];
  if (dinv && threadIdx.x < bb) dinv[k * bb + threadIdx.x] = lu[((size_t)r * n + r) * bb + threadIdx.x];
}

template <bool BACKWARD>
__device__ __forceinline__ void lu_sweep(const LuProg P, const double *__restrict__ rhs /* forward: d[row]; backward: dinv[k] */, double *vs, double *prod)
{
  const int lane = threadIdx.x;
  if (P.n == 0) return;
  int row_n = P.row[0], o_n = P.ptr[0], e_n = P.ptr[1];
  double r_n = BACKWARD ? rhs[0] : rhs[row_n];
  double nv[LUS_PRE]; int nc[LUS_PRE];
#pragma unroll
  for (int q = 0; q < LUS_PRE; q++) { const int e = o_n + lane + 32 * q; nv[q] = e < e_n ? P.val[e] : 0.0; nc[q] = e < e_n ? P.col[e] : 0; }
  for (int k = 0; k < P.n; k++) {
    const int row = row_n, o = o_n, cnt = e_n - o_n;
    const double r = r_n;
    double cv[LUS_PRE]; int cc[LUS_PRE];
#pragma unroll
    for (int q = 0; q < LUS_PRE; q++) { cv[q] = nv[q]; cc[q] = nc[q]; }
    if (k + 1 < P.n) {                       // the next row's entries travel while this row is summed
      row_n = P.row[k + 1]; o_n = e_n; e_n = P.ptr[k + 2];
      r_n = BACKWARD ? rhs[k + 1] : rhs[row_n];
#pragma unroll
      for (int q = 0; q < LUS_PRE; q++) { const int e = o_n + lane + 32 * q; nv[q] = e < e_n ? P.val[e] : 0.0; nc[q] = e < e_n ? P.col[e] : 0; }
    }
#pragma unroll
    for (int q = 0; q < LUS_PRE; q++) { const int kk = lane + 32 * q; if (kk < cnt) prod[kk] = cv[q] * vs[cc[q]]; }
    for (int kk = lane + 32 * LUS_PRE; kk < cnt; kk += 32) prod[kk] = P.val[o + kk] * vs[P.col[o + kk]];
    __syncwarp();
    if (lane == 0) {
      double s = 0.0;
#pragma unroll 8
      for (int kk = 0; kk < cnt; kk++) s += prod[kk];
      vs[row] = BACKWARD ? (vs[row] - s) * r : r - s;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32) k_lu_solve_lists(int N, LuProg F, LuProg B, const double *__restrict__ dinv, double *__restrict__ v, const double *__restrict__ d)
{
  __shared__ double vs[LU_MAX_N];
  __shared__ double prod[LU_MAX_N];
  for (int i = threadIdx.x; i < N; i += 32) vs[i] = 0.0;      // rows with VCLASS < ACTIVE_CLASS stay 0 (ugiter.cc:4488)
  __syncwarp();
  lu_sweep<false>(F, d, vs, prod);
  lu_sweep<true>(B, dinv, vs, prod);
  for (int i = threadIdx.x; i < N; i += 32) v[i] = vs[i];
}

// Block rows: l_luiter ugiter.cc:4522-4795.  Per entry the lanes form e_i = (m_i0 w_0 + m_i1 w_1) + m_i2 w_2 (MATMUL_nn, ugblas.h:161-213),
// lane 0 adds them in list order; forward v = d - sum (Diag(L) = I), backward v = Inv * (v - sum) (SolveInverseSmallBlock block.cc:225).
template <int BS, bool BACKWARD>
__device__ __forceinline__ void lu_sweep_block(const LuProg P, const double *__restrict__ d, const double *__restrict__ dinv, double *vs, double *prod)
{
  constexpr int BB = BS * BS;
  const int lane = threadIdx.x;
  for (int k = 0; k < P.n; k++) {
    const int row = P.row[k], o = P.ptr[k], cnt = P.ptr[k + 1] - o;
    for (int kk = lane; kk < cnt; kk += 32) {
      const double *m = P.val + (size_t)(o + kk) * BB, *w = vs + (size_t)P.col[o + kk] * BS;
#pragma unroll
      for (int i = 0; i < BS; i++) {
        double t = m[i * BS] * w[0];
#pragma unroll
        for (int q = 1; q < BS; q++) t = t + m[i * BS + q] * w[q];
        prod[kk * BS + i] = t;
      }
    }
    __syncwarp();
    if (lane == 0) {
      double acc[BS], s[BS];
#pragma unroll
      for (int i = 0; i < BS; i++) acc[i] = 0.0;
      for (int kk = 0; kk < cnt; kk++) {
#pragma unroll
        for (int i = 0; i < BS; i++) acc[i] += prod[kk * BS + i];
      }
      if (BACKWARD) {
#pragma unroll
        for (int i = 0; i < BS; i++) s[i] = vs[row * BS + i] - acc[i];
        const double *inv = dinv + (size_t)k * BB;
#pragma unroll
        for (int i = 0; i < BS; i++) {
          double sum = 0.0;
#pragma unroll
          for (int j = 0; j < BS; j++) sum += inv[i * BS + j] * s[j];
          vs[row * BS + i] = sum;
        }
      } else {
#pragma unroll
        for (int i = 0; i < BS; i++) vs[row * BS + i] = d[row * BS + i] - acc[i];
      }
    }
    __syncwarp();
  }
}

template <int BS>
__global__ void __launch_bounds__(32) k_lu_solve_lists_block(int N, LuProg F, LuProg B, const double *__restrict__ dinv, double *__restrict__ v, const double *__restrict__ d)
{
  __shared__ double vs[LU_MAX_N];
  __shared__ double prod[LU_MAX_N];
  for (int i = threadIdx.x; i < N; i += 32) vs[i] = 0.0;
  __syncwarp();
  lu_sweep_block<BS, false>(F, d, dinv, vs, prod);
  lu_sweep_block<BS, true>(B, d, dinv, vs, prod);
  for (int i = threadIdx.x; i < N; i += 32) v[i] = vs[i];
}

// ---- triangular sweeps with all warps: every row adds ITS terms in list order, rows overlap as far as their dependencies allow ----------
// The one-warp sweeps above finish a row before they start the next: 2 x (active rows) x (entries per row) dependent additions, 0.9 ms
// for the 567-row base level of the 1025 x 1025 x 769 hierarchy.  A row's sum is sequential by definition (the reference's order), but
// different rows are independent up to the values they read.  Warp w takes the program rows w, w + W, ...; a row's list is walked in
// chunks: the lanes wait TOGETHER until the rows their 32 entries read are published (an epoch word per row in shared memory), form
// the products, lane 0 adds them in list order.  The last chunk holds only the last LUW_TAIL entries: the entry that reads the preceding
// row sits near the END of a list (original connections follow the fill-in, gm/algebra.cc:1051-1078), so almost all of a row's sum is
// formed while its predecessors are still busy, and what remains on the critical path is a hand-over and a few additions per row.
// (Measured and rejected: one THREAD per row with per-lane waits -- lanes of a warp waiting for one another: 4.5 x slower than one warp.)
// Progress: dependencies point to earlier program rows and every warp takes its rows in program order, so the earliest unfinished row
// never waits; a bounded spin reports through the error word instead of hanging.
#define LUW_TAIL 8
#define LUW_SPIN_MAX (1 << 22)
template <int BS, bool BACKWARD>
__device__ __forceinline__ void lu_sweep_warps(const LuProg P, const double *rhs /* forward: d */, const double *__restrict__ dinv, volatile double *vs, volatile int *flag,
                                               int epoch, double *pbuf_all, int *err)
{
  constexpr int BB = BS * BS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  volatile double *pbuf = pbuf_all + (size_t)warp * 32 * BS;
  for (int k = warp; k < P.n; k += nw) {
    const int row = P.row[k], o = P.ptr[k], cnt = P.ptr[k + 1] - o;
    const int head = cnt > LUW_TAIL ? cnt - LUW_TAIL : 0;
    // what the END of the row needs is requested first: the tail chunk's entries, the right-hand side / inverse diagonal
    const bool thave = head + lane < cnt;
    const int tc = thave ? P.col[o + head + lane] : 0;
    double tm[BB], fin[BB];
    if (thave) {
#pragma unroll
      for (int q = 0; q < BB; q++) tm[q] = P.val[(size_t)(o + head + lane) * BB + q];
    }
    if (lane == 0) {
      if (BACKWARD) {
#pragma unroll
        for (int q = 0; q < BB; q++) fin[q] = dinv[(size_t)k * BB + q];
      } else {
#pragma unroll
        for (int i = 0; i < BS; i++) fin[i] = rhs[row * BS + i];
      }
    }
    double acc[BS];
#pragma unroll
    for (int i = 0; i < BS; i++) acc[i] = 0.0;
    int a = 0;
    while (a < cnt) {
      const bool tail = a >= head;
      const int b = tail ? cnt : (a + 32 < head ? a + 32 : head);
      const bool have = a + lane < b;
      const int e = o + a + lane;
      const int c = tail ? tc : (have ? P.col[e] : 0);
      double m[BB];
      if (tail) {
#pragma unroll
        for (int q = 0; q < BB; q++) m[q] = tm[q];
      } else if (have) {
#pragma unroll
        for (int q = 0; q < BB; q++) m[q] = P.val[(size_t)e * BB + q];
      }
      for (int it = 0; !__all_sync(0xffffffffu, !have || flag[c] == epoch); it++)
        if (it > LUW_SPIN_MAX) { if (lane == 0) atomicExch(err, UGGPU_ERROR); break; }
      __threadfence_block();
      if (have) {
        if (BS == 1) pbuf[lane] = m[0] * vs[c];
        else {
#pragma unroll
          for (int i = 0; i < BS; i++) {
            double t = m[i * BS] * vs[c * BS];
#pragma unroll
            for (int q = 1; q < BS; q++) t = t + m[i * BS + q] * vs[c * BS + q];
            pbuf[lane * BS + i] = t;
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        for (int kk = 0; kk < b - a; kk++) {
#pragma unroll
          for (int i = 0; i < BS; i++) acc[i] += pbuf[kk * BS + i];
        }
      }
      __syncwarp();
      a = b;
    }
    if (lane == 0) {
      if (BACKWARD) {
        if (BS == 1) vs[row] = (vs[row] - acc[0]) * fin[0];
        else {
          double sv[BS], out[BS];
#pragma unroll
          for (int i = 0; i < BS; i++) sv[i] = vs[row * BS + i] - acc[i];
#pragma unroll
          for (int i = 0; i < BS; i++) {
            double sum = 0.0;
#pragma unroll
            for (int j = 0; j < BS; j++) sum += fin[i * BS + j] * sv[j];
            out[i] = sum;
          }
#pragma unroll
          for (int i = 0; i < BS; i++) vs[row * BS + i] = out[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < BS; i++) vs[row * BS + i] = fin[i] - acc[i];
      }
      __threadfence_block();
      flag[row] = epoch;
    }
    __syncwarp();
  }
}

// ---- the whole base-level solve in ONE kernel -------------------------------------------------------------------------------------
// `ls $I lu` around l_luiter: LinearResiduum, then per iteration v = (LU)^-1 b; b -= A v; c += v; LinearResiduum; convergence test
// (ls.cc:637-749).  One CTA: warp 0 runs the list-ordered sweeps above, all threads the defect update (one thread per row, the row's
// terms in VSTART->MNEXT order like k_dmatmul_k) and the norms.  The defect lives in shared memory for the duration.  No host round
// trip: the cycle's stream is never drained in the middle (the host-driven loop base_solve_host read two norms back per cycle).
struct BaseArgs { int N, n, maxit; double abslimit, reduction; int par; };

template <int BS>
__device__ __forceinline__ void base_norm(const double *bsh, const uint8_t *__restrict__ ctl, int n, double *red, double *out)
{
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  double acc[BS];
#pragma unroll
  for (int i = 0; i < BS; i++) acc[i] = 0.0;
  for (int r = tid; r < n; r += blockDim.x)
    if (ctl[r] & UGGPU_CTL_NEW_DEFECT) {
#pragma unroll
      for (int i = 0; i < BS; i++) acc[i] += bsh[r * BS + i] * bsh[r * BS + i];
    }
#pragma unroll
  for (int i = 0; i < BS; i++) {
    double v = acc[i];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[w * BS + i] = v;
  }
  __syncthreads();
  if (tid < BS) {
    double v = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) v += red[k * BS + tid];
    out[tid] = sqrt(v);
  }
  __syncthreads();
}

template <int BS>
__global__ void __launch_bounds__(LU_THREADS) k_base_solve(SellView A, LuProg F, LuProg B, const double *__restrict__ dinv, const uint8_t *__restrict__ ctl,
                                                           double *__restrict__ c, double *__restrict__ b, BaseArgs a, int *err)
{
  extern __shared__ double smem_base[];
  double *vs = smem_base, *prod = vs + LU_MAX_N, *bsh = prod + LU_MAX_N, *red = bsh + LU_MAX_N, *nrm = red + 32 * BS, *reach = nrm + BS;
  int *flag = reinterpret_cast<int *>(prod);             // all-warp sweeps: one epoch word per row (the one-warp sweeps use the array for products)
  double *pbuf = reach + BS;                             // ... and 32 * BS products per warp
  __shared__ int stop;
  constexpr int BB = BS * BS;
  const int tid = threadIdx.x, N = a.N;
  for (int i = tid; i < N; i += blockDim.x) bsh[i] = b[i];
  __syncthreads();
  base_norm<BS>(bsh, ctl, a.n, red, nrm);
  if (tid == 0) {
    bool below = true;                                   // sc_cmp(first, abslimit) npscan.cc:1027
    for (int i = 0; i < BS; i++) {
      if (fabs(nrm[i]) >= fabs(a.abslimit)) below = false;
      reach[i] = nrm[i] * a.reduction;
      if (reach[i] == 0.0) reach[i] = a.reduction;       // ls.cc:663-667
    }
    stop = below ? 1 : 0;
  }
  __syncthreads();
  if (stop) return;
  for (int it = 0; it < a.maxit; it++) {
    if (a.par) {
      for (int i = tid; i < N; i += blockDim.x) vs[i] = 0.0;     // rows with VCLASS < ACTIVE_CLASS stay 0 (ugiter.cc:4488)
      if (it == 0) for (int i = tid; i < a.n; i += blockDim.x) flag[i] = 0;
      __syncthreads();
      lu_sweep_warps<BS, false>(F, bsh, dinv, vs, flag, 2 * it + 1, pbuf, err);
      __syncthreads();
      lu_sweep_warps<BS, true>(B, bsh, dinv, vs, flag, 2 * it + 2, pbuf, err);
    } else if (tid < 32) {
      for (int i = tid; i < N; i += 32) vs[i] = 0.0;
      __syncwarp();
      if (BS == 1) { lu_sweep<false>(F, bsh, vs, prod); lu_sweep<true>(B, dinv, vs, prod); }
      else { lu_sweep_block<BS, false>(F, bsh, dinv, vs, prod); lu_sweep_block<BS, true>(B, bsh, dinv, vs, prod); }
    }
    __syncthreads();
    // b -= A v (dmatmul_minus, all rows), c += v
    for (int r = tid; (r & ~31) < a.n; r += blockDim.x) {
      if (r < a.n) {
        const int lane = r & 31;
        const int64_t sp = slice_off(A, r >> 5);
        const int64_t cpo = (A.fixed_w && A.col_ptr == A.slice_ptr) ? sp : A.col_ptr[r >> 5];
        const int len = (int)A.rowlen[r];
        const double *tp = (cpo < 0 && A.vt && UG_VALTAB(cpo) >= 0) ? A.vt + UG_VALTAB(cpo) : nullptr;
        const double *vp = A.val + sp * BB + lane;
        const ColIter ci = col_iter(A, r);
        double sm[BS];
#pragma unroll
        for (int i = 0; i < BS; i++) sm[i] = 0.0;
        for (int j = 0; j < len; j++) {
          const int cc = col_at(ci, j);
          double m[BB];
#pragma unroll
          for (int k = 0; k < BB; k++) m[k] = tp ? tp[(size_t)j * BB + k] : vp[((size_t)j * BB + k) * 32];
#pragma unroll
          for (int i = 0; i < BS; i++) {
            double acc = m[i * BS] * vs[cc * BS];
#pragma unroll
            for (int q = 1; q < BS; q++) acc = acc + m[i * BS + q] * vs[cc * BS + q];
            sm[i] += acc;
          }
        }
#pragma unroll
        for (int i = 0; i < BS; i++) bsh[r * BS + i] = bsh[r * BS + i] - sm[i];
      }
    }
    for (int i = tid; i < N; i += blockDim.x) c[i] = c[i] + vs[i];
    __syncthreads();
    base_norm<BS>(bsh, ctl, a.n, red, nrm);
    if (tid == 0) {
      bool b1 = true, b2 = true;
      for (int i = 0; i < BS; i++) { if (fabs(nrm[i]) >= fabs(a.abslimit)) b1 = false; if (fabs(nrm[i]) >= fabs(reach[i])) b2 = false; }
      stop = (b1 || b2) ? 1 : 0;
    }
    __syncthreads();
    if (stop) break;
  }
  for (int i = tid; i < N; i += blockDim.x) b[i] = bsh[i];
}

int level_free_lu(uggpu_ctx *ctx, Level *L)
{
  const size_t bbf = (size_t)(L->bs > 0 ? L->bs * L->bs : 1), nrow = (size_t)(L->bs > 0 ? L->luN / L->bs : L->luN);
  if (L->lu) dfree(ctx, L->lu, (size_t)L->luN * L->luN);
  if (L->lu_lo_ptr) dfree(ctx, L->lu_lo_ptr, nrow + 1);
  if (L->lu_up_ptr) dfree(ctx, L->lu_up_ptr, nrow + 1);
  if (L->lu_lo_row) dfree(ctx, L->lu_lo_row, nrow + 1);
  if (L->lu_up_row) dfree(ctx, L->lu_up_row, nrow + 1);
  if (L->lu_dinv) dfree(ctx, L->lu_dinv, (nrow + 1) * bbf);
  if (L->lu_lo_col) dfree(ctx, L->lu_lo_col, (size_t)L->lu_lo_nnz + 1);
  if (L->lu_up_col) dfree(ctx, L->lu_up_col, (size_t)L->lu_up_nnz + 1);
  if (L->lu_lo_val) dfree(ctx, L->lu_lo_val, ((size_t)L->lu_lo_nnz + 1) * bbf);
  if (L->lu_up_val) dfree(ctx, L->lu_up_val, ((size_t)L->lu_up_nnz + 1) * bbf);
  L->luN = 0; L->luA = -1; L->luGen = -1; L->lu_lo_nnz = L->lu_up_nnz = L->lu_active = 0;
  return 0;
}

// Replays the list operations of l_lrdecomp (scalar path ugiter.cc:3715-3768) on the pattern of M and the finished factors:
// row lists start in VSTART->MNEXT order; at step i, for every list entry j > i whose multiplier L_ji is not zero (:3741) and
// every list entry k > i, a missing connection (j,k) is inserted at the second place of the lists of j and of k.
static int lu_lists(uggpu_ctx *ctx, Level *L, const SellMat *M)
{
  const int n = L->n;
  std::vector<int32_t> rowptr((size_t)n + 1), col((size_t)M->nnz);
  UG_TRY(sell_to_host_csr(ctx, M, rowptr.data(), col.data(), nullptr));
  std::vector<uint8_t> vclass((size_t)n);
  const int bs = L->bs, bb = bs * bs;
  std::vector<double> lu((size_t)n * n * bb);
  CUDA_TRY(cudaMemcpyAsync(vclass.data(), L->vclass, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(lu.data(), L->lu, sizeof(double) * (size_t)n * n * bb, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // singly linked row lists in one node pool: node = (column, next); head[r] = the diagonal entry
  std::vector<int32_t> ncol, nnext, head((size_t)n, -1);
  ncol.reserve(col.size() * 2); nnext.reserve(col.size() * 2);
  std::vector<uint8_t> present((size_t)n * n, 0);
  for (int r = 0; r < n; r++) {
    int prev = -1;
    for (int e = rowptr[r]; e < rowptr[r + 1]; e++) {
      const int id = (int)ncol.size();
      ncol.push_back(col[e]); nnext.push_back(-1);
      if (prev < 0) head[r] = id; else nnext[prev] = id;
      prev = id;
      present[(size_t)r * n + col[e]] = 1;
    }
    if (head[r] < 0 || ncol[head[r]] != r) return uggpu_fail(UGGPU_ERROR, "base level: row %d does not start with its diagonal entry", r);
  }
  auto active = [&](int x) { return vclass[x] >= 3; };
  auto insert_second = [&](int r, int c) {
    const int id = (int)ncol.size();
    ncol.push_back(c); nnext.push_back(nnext[head[r]]);
    nnext[head[r]] = id;
    present[(size_t)r * n + c] = 1;
  };
  for (int i = 0; i < n; i++) {
    if (!active(i)) continue;
    for (int a = nnext[head[i]]; a >= 0; a = nnext[a]) {
      const int j = ncol[a];
      if (!(active(j) && j > i)) continue;
      const double *lji = &lu[((size_t)i * n + j) * bb];    // multiplier block L_ji (block (j,i) at lu[(i*n+j)*bb])
      bool pivzero = true;
      for (int q = 0; q < bb; q++) if (lji[q] != 0.0) pivzero = false;
      if (pivzero) continue;                                  // ugiter.cc:3741 / :3829
      for (int c = nnext[head[i]]; c >= 0; c = nnext[c]) {
        const int k = ncol[c];
        if (!(active(k) && k > i)) continue;
        if (present[(size_t)j * n + k]) continue;
        if (bs > 1) {                                         // block path only: a zero correction creates nothing (:3867)
          double cor[UGGPU_MAX_BS * UGGPU_MAX_BS];
          const double *uik = &lu[((size_t)k * n + i) * bb];
          if (bs == 2 ? block_mul<2>(lji, uik, cor) : block_mul<3>(lji, uik, cor)) continue;
        }
        insert_second(j, k);
        insert_second(k, j);
      }
    }
  }
  // row programs: active rows ascending with their lower entries, descending with their upper entries (list order)
  std::vector<int32_t> f_row, f_ptr(1, 0), f_col, b_row, b_ptr(1, 0), b_col;
  for (int r = 0; r < n; r++) {
    if (!active(r)) continue;
    f_row.push_back(r);
    for (int a = nnext[head[r]]; a >= 0; a = nnext[a]) { const int c = ncol[a]; if (active(c) && c < r) f_col.push_back(c); }
    f_ptr.push_back((int32_t)f_col.size());
  }
  for (int r = n - 1; r >= 0; r--) {
    if (!active(r)) continue;
    b_row.push_back(r);
    for (int a = nnext[head[r]]; a >= 0; a = nnext[a]) { const int c = ncol[a]; if (active(c) && c > r) b_col.push_back(c); }
    b_ptr.push_back((int32_t)b_col.size());
  }
  const int na = (int)f_row.size();
  L->lu_active = na; L->lu_lo_nnz = (int)f_col.size(); L->lu_up_nnz = (int)b_col.size();
  UG_TRY(dalloc(ctx, &L->lu_lo_row, (size_t)n + 1)); UG_TRY(dalloc(ctx, &L->lu_up_row, (size_t)n + 1));
  UG_TRY(dalloc(ctx, &L->lu_lo_ptr, (size_t)n + 1)); UG_TRY(dalloc(ctx, &L->lu_up_ptr, (size_t)n + 1));
  UG_TRY(dalloc(ctx, &L->lu_dinv, ((size_t)n + 1) * bb));
  UG_TRY(dalloc(ctx, &L->lu_lo_col, (size_t)L->lu_lo_nnz + 1)); UG_TRY(dalloc(ctx, &L->lu_up_col, (size_t)L->lu_up_nnz + 1));
  UG_TRY(dalloc(ctx, &L->lu_lo_val, ((size_t)L->lu_lo_nnz + 1) * bb)); UG_TRY(dalloc(ctx, &L->lu_up_val, ((size_t)L->lu_up_nnz + 1) * bb));
  cudaStream_t st = ctx->stream;
  if (na) {
    CUDA_TRY(cudaMemcpyAsync(L->lu_lo_row, f_row.data(), sizeof(int32_t) * na, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(L->lu_up_row, b_row.data(), sizeof(int32_t) * na, cudaMemcpyHostToDevice, st));
  }
  CUDA_TRY(cudaMemcpyAsync(L->lu_lo_ptr, f_ptr.data(), sizeof(int32_t) * (na + 1), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(L->lu_up_ptr, b_ptr.data(), sizeof(int32_t) * (na + 1), cudaMemcpyHostToDevice, st));
  if (L->lu_lo_nnz) CUDA_TRY(cudaMemcpyAsync(L->lu_lo_col, f_col.data(), sizeof(int32_t) * f_col.size(), cudaMemcpyHostToDevice, st));
  if (L->lu_up_nnz) CUDA_TRY(cudaMemcpyAsync(L->lu_up_col, b_col.data(), sizeof(int32_t) * b_col.size(), cudaMemcpyHostToDevice, st));
  if (na) {
    k_lu_pack<<<na, 128, 0, st>>>(n, bb, L->lu, L->lu_lo_row, L->lu_lo_ptr, L->lu_lo_col, L->lu_lo_val, nullptr);
    KCHECK(ctx);
    k_lu_pack<<<na, 128, 0, st>>>(n, bb, L->lu, L->lu_up_row, L->lu_up_ptr, L->lu_up_col, L->lu_up_val, L->lu_dinv);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaStreamSynchronize(st));      // the host vectors go out of scope
  return 0;
}

static int ensure_vec(uggpu_ctx *ctx, int level, int vec) { return uggpu_vec_alloc(ctx, level, vec); }

extern "C" int uggpu_lmgc_preprocess(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int A)
{
  if (!ctx || !cfg) return uggpu_fail(UGGPU_ERROR, "null argument");
  int bl = cfg->baselevel;
  if (bl > level) return uggpu_fail(UGGPU_ERROR, "baselevel %d above level %d", bl, level);
  for (int l = bl; l <= level; l++) {
    if (!get_level(ctx, l)) return UGGPU_ERROR;
    if (!get_mat(ctx, l, A)) return UGGPU_DESC_MISMATCH;
    if (l > bl && (!ctx->lev[l].P.valid() || !ctx->lev[l].R.valid()))
      return uggpu_fail(UGGPU_NO_COARSER_GRID, "level %d has no transfer stencils", l);
    UG_TRY(ensure_vec(ctx, l, cfg->t));
    UG_TRY(ensure_vec(ctx, l, UGGPU_VEC_TMP_B));
    if (l > bl && cfg->smoother != UGGPU_SM_JAC) {          // GSPreProcess / SGSPreProcess / SORPreProcess (iter.cc:1003,1353,4717)
      if (cfg->smoother < UGGPU_SM_JAC || cfg->smoother > UGGPU_SM_ILU) return uggpu_fail(UGGPU_ERROR, "lmgc: unknown smoother class %d", cfg->smoother);
      if (cfg->smoother == UGGPU_SM_ILU) {               // ILUPreProcess iter.cc:5444: L = copy of A, decomposed in place
        if (cfg->smoother_L == A) return uggpu_fail(UGGPU_DESC_MISMATCH, "lmgc: the ILU decomposition needs its own matrix handle");
        UG_TRY(uggpu_dmatcopy(ctx, l, l, UGGPU_ALL_VECTORS, cfg->smoother_L, A));
        UG_TRY(uggpu_l_ilubthdecomp(ctx, l, cfg->smoother_L, cfg->ilu_beta));
        continue;
      }
      UG_TRY(uggpu_gs_preprocess(ctx, l, A));
      if (cfg->smoother == UGGPU_SM_SGS) UG_TRY(ensure_vec(ctx, l, UGGPU_VEC_TMP_A));
    }
  }
  if (cfg->base_solver == nullptr) {
    Level *L = &ctx->lev[bl];
    int N = L->n * L->bs;
    if (N > LU_MAX_N) return uggpu_fail(UGGPU_OUT_OF_MEM, "base level has %d unknowns; the device LU handles at most %d (use a coarser base level or a host base solver)", N, LU_MAX_N);
    UG_TRY(level_free_lu(ctx, L));
    L->luN = N; L->luA = A;
    { SellMat *M0 = get_mat(ctx, bl, A); L->luGen = M0 ? M0->gen : -1; }
    UG_TRY(dalloc(ctx, &L->lu, (size_t)N * N));
    if (N > 0) {
      CUDA_TRY(cudaMemsetAsync(L->lu, 0, (size_t)N * N * sizeof(double), ctx->stream));
      SellMat *M = get_mat(ctx, bl, A);
      const bool lists = !getenv("UGGPU_LU_INDEX_ORDER");     // A/B switch: the first implementation (scalar elimination, sums in index order)
      if (L->bs > 1 && lists) {
        // block rows: the reference's block elimination on bs x bs blocks (n*n blocks of bb doubles = N*N doubles)
        k_lu_scatter_block<<<(L->n + 255) / 256, 256, 0, ctx->stream>>>(view(*M), L->bs, L->n, L->lu);
        KCHECK(ctx);
        if (L->bs == 2) k_lu_factor_block<2><<<1, LU_THREADS, 0, ctx->stream>>>(L->n, L->vclass, L->lu, ctx->derr);
        else k_lu_factor_block<3><<<1, LU_THREADS, 0, ctx->stream>>>(L->n, L->vclass, L->lu, ctx->derr);
        KCHECK(ctx);
      } else {
        k_lu_scatter<<<(L->n + 255) / 256, 256, 0, ctx->stream>>>(view(*M), L->bs, N, L->lu);
        KCHECK(ctx);
        k_lu_factor<<<1, LU_THREADS, 0, ctx->stream>>>(N, L->bs, L->vclass, L->lu);
        KCHECK(ctx);
      }
      // the summation order of the reference's matrix lists with fill-in (base level)
      if (lists) UG_TRY(lu_lists(ctx, L, M));
    }
    UG_TRY(ensure_vec(ctx, bl, UGGPU_VEC_TMP_C));
  }
  return 0;
}

static int sc_cmp(const double *x, const double *y, int n)   // npscan.cc:1027
{
  for (int i = 0; i < n; i++) if (fabs(x[i]) >= fabs(y[i])) return 0;
  return 1;
}

// dnrm2x of one level (rowmode) -> out[bs]
static int level_norm(uggpu_ctx *ctx, int level, int rowmode, const double *x, double *out)
{
  int bs = ctx->lev[level].bs;
  UG_TRY(k_reduce(ctx, level, rowmode, RED_NRM2, x, x, 0));
  UG_TRY(fetch_results(ctx, 1));
  for (int i = 0; i < bs; i++) out[i] = sqrt(ctx->hres[i]);
  return 0;
}

// base level: NP_LINEAR_SOLVER `ls $I lu` = LinearResiduum + LinearSolver (ls.cc:577,637) with Iter = LU (damp 1)
static int base_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A)
{
  if (cfg->base_solver) {
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    int rc = cfg->base_solver(cfg->base_user, ctx, level, c, b, A);
    if (rc) return uggpu_fail(rc, "host base solver returned %d", rc);
    return 0;
  }
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  if (!L->lu || L->luA != A) return uggpu_fail(UGGPU_ERROR, "base level %d not factored: call uggpu_lmgc_preprocess first", level);
  {
    SellMat *M0 = get_mat(ctx, level, A);
    if (M0 && M0->gen != L->luGen)
      return uggpu_fail(UGGPU_ERROR, "base level %d: the matrix values changed after the factorisation (uggpu_mat_set_values / uggpu_galerkin / uggpu_dmatcopy): "
                                     "call uggpu_lmgc_preprocess again", level);
  }
  int bs = L->bs;
  double *cp = get_vec(ctx, level, c), *bp = get_vec(ctx, level, b), *cc = get_vec(ctx, level, UGGPU_VEC_TMP_C);
  if (!cp || !bp || !cc) return UGGPU_DESC_MISMATCH;
  if (L->n == 0) return 0;
  // the whole loop on the device (one CTA) when the level is held completely and the list-ordered factors exist
  if (L->lu_lo_ptr && !(ctx->comm && L->partitioned) && !getenv("UGGPU_BASE_HOST_LOOP")) {
    const LuProg F{L->lu_lo_row, L->lu_lo_ptr, L->lu_lo_col, L->lu_lo_val, L->lu_active}, B{L->lu_up_row, L->lu_up_ptr, L->lu_up_col, L->lu_up_val, L->lu_active};
    const BaseArgs ba{L->luN, L->n, cfg->base_maxit, cfg->base_abslimit, cfg->base_reduction, getenv("UGGPU_LU_ONE_WARP") ? 0 : 1};
    SellMat *M = get_mat(ctx, level, A);
    if (!M) return UGGPU_DESC_MISMATCH;
    const size_t smem = sizeof(double) * (3 * LU_MAX_N + 32 * UGGPU_MAX_BS + 2 * UGGPU_MAX_BS + (LU_THREADS / 32) * 32 * UGGPU_MAX_BS);
    // per device and cheap: set on every call (a process may drive several devices)
    if (bs == 1) CUDA_TRY(cudaFuncSetAttribute(k_base_solve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else if (bs == 2) CUDA_TRY(cudaFuncSetAttribute(k_base_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CUDA_TRY(cudaFuncSetAttribute(k_base_solve<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, UGGPU_K_BASE, level, 8.0 * L->luN * L->luN);
    if (bs == 1) k_base_solve<1><<<1, LU_THREADS, smem, ctx->stream>>>(view(*M), F, B, L->lu_dinv, L->ctl, cp, bp, ba, ctx->derr);
    else if (bs == 2) k_base_solve<2><<<1, LU_THREADS, smem, ctx->stream>>>(view(*M), F, B, L->lu_dinv, L->ctl, cp, bp, ba, ctx->derr);
    else k_base_solve<3><<<1, LU_THREADS, smem, ctx->stream>>>(view(*M), F, B, L->lu_dinv, L->ctl, cp, bp, ba, ctx->derr);
    KCHECK(ctx);
    return 0;
  }
  double first[UGGPU_MAX_BS], last[UGGPU_MAX_BS], reach[UGGPU_MAX_BS], absl[UGGPU_MAX_BS];
  UG_TRY(level_norm(ctx, level, 1, bp, last));
  for (int i = 0; i < bs; i++) {
    first[i] = last[i]; absl[i] = cfg->base_abslimit;
    reach[i] = first[i] * cfg->base_reduction;
    if (reach[i] == 0.0) reach[i] = cfg->base_reduction;
  }
  if (sc_cmp(first, absl, bs)) return 0;
  Damp none = mkdamp(nullptr, 0);
  for (int it = 0; it < cfg->base_maxit; it++) {
    {
      ProfScope ps(ctx, UGGPU_K_BASE, level, 8.0 * L->luN * L->luN);
      if (L->lu_lo_ptr) {
        const LuProg F{L->lu_lo_row, L->lu_lo_ptr, L->lu_lo_col, L->lu_lo_val, L->lu_active}, B{L->lu_up_row, L->lu_up_ptr, L->lu_up_col, L->lu_up_val, L->lu_active};
        if (bs == 1) k_lu_solve_lists<<<1, 32, 0, ctx->stream>>>(L->luN, F, B, L->lu_dinv, cc, bp);
        else if (bs == 2) k_lu_solve_lists_block<2><<<1, 32, 0, ctx->stream>>>(L->luN, F, B, L->lu_dinv, cc, bp);
        else k_lu_solve_lists_block<3><<<1, 32, 0, ctx->stream>>>(L->luN, F, B, L->lu_dinv, cc, bp);
      }
      else k_lu_solve<<<1, LU_THREADS, 0, ctx->stream>>>(L->luN, bs, L->vclass, L->lu, cc, bp);
      KCHECK(ctx);
    }
    UG_TRY(k_dmatmul(ctx, level, 2, 0, b, A, UGGPU_VEC_TMP_C));
    UG_TRY(k_vec_op(ctx, level, 0, VOP_ADD, cp, cc, none));
    UG_TRY(level_norm(ctx, level, 1, bp, last));
    if (sc_cmp(last, absl, bs) || sc_cmp(last, reach, bs)) break;
  }
  return 0;
}

// ---- Lmgc ---------------------------------------------------------------------------------------------------------------
struct TopFuse {       // work of the enclosing LinearSolver iteration folded into the cycle's kernels (fused schedule only)
  int level = -1;      // solver level; -1: none
  int x = 0;           // x += c (LSUpdate ls.cc:869) on `level`
  bool c_zero = false; // c is known to be 0 on entry (dset ls.cc:695 skipped)
  bool norm = false;   // accumulate ||b||^2 over NEW_DEFECT rows of `level` into result slot 0
  bool done_x = false, done_norm = false;
  // another cycle follows (LinearSolver's loop): the last smoothing step also writes the damped Jacobi correction of the NEW defect, which
  // is what the next cycle starts with (l_jac + dscalx of its first smoothing step) -- the Jacobi pass over the top level at the start
  // of every cycle but the first disappears.  next_t: where it was written (0: not done, 1: cfg->t, 2: the second temporary)
  bool want_t = false;
  int next_t = 0;
};

static int lmgc_unfused(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A)
{
  if (level <= cfg->baselevel) return base_solve(ctx, cfg, level, c, b, A);
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  const int t = cfg->t;
  double one[UGGPU_MAX_BS] = {1.0, 1.0, 1.0};
  for (int i = 0; i < cfg->nu1; i++) {
    UG_TRY(uggpu_smooth(ctx, level, cfg->smoother, t, b, A, cfg->smooth_damp, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : UGGPU_VEC_TMP_A));
    UG_TRY(uggpu_dadd(ctx, level, level, UGGPU_ALL_VECTORS, c, t));
  }
  UG_TRY(uggpu_restrict(ctx, level, b, b, one));                                   // iter.cc:7843, Factor_One
  UG_TRY(uggpu_dset(ctx, level - 1, level - 1, UGGPU_ALL_VECTORS, c, 0.0));        // :7873
  for (int g = 0; g < cfg->gamma; g++) UG_TRY(lmgc_unfused(ctx, cfg, level - 1, c, b, A));
  UG_TRY(uggpu_interpolate_correction(ctx, level, t, c, cfg->cycle_damp));         // :7886
  UG_TRY(uggpu_dadd(ctx, level, level, UGGPU_ALL_VECTORS, c, t));                  // :7903
  UG_TRY(uggpu_dmatmul_minus(ctx, level, level, UGGPU_ALL_VECTORS, b, A, t));      // :7905
  for (int i = 0; i < cfg->nu2; i++) {
    UG_TRY(uggpu_smooth(ctx, level, cfg->smoother, t, b, A, cfg->smooth_damp, cfg->smoother == UGGPU_SM_ILU ? cfg->smoother_L : UGGPU_VEC_TMP_A));
    UG_TRY(uggpu_dadd(ctx, level, level, UGGPU_ALL_VECTORS, c, t));
  }
  if (cfg->level_opt) UG_TRY(uggpu_minimize_level(ctx, level, c, b, A, t));         // :7944 AdaptCorrection (transfer $L; np->t is free again, :7942)
  return 0;
}

// MinimizeLevel np/procs/transfer.cc:488-516, call for call
extern "C" int uggpu_minimize_level(uggpu_ctx *ctx, int level, int c, int b, int A, int t)
{
  double a0 = 0.0, a1 = 0.0;
  UG_TRY(uggpu_vec_alloc(ctx, level, t));
  UG_TRY(uggpu_dmatmul(ctx, level, level, UGGPU_ALL_VECTORS, t, A, c));             // :498
  UG_TRY(uggpu_ddot(ctx, level, level, UGGPU_ALL_VECTORS, t, b, &a0));              // :504
  UG_TRY(uggpu_dnrm2(ctx, level, level, UGGPU_ALL_VECTORS, t, &a1));                // :506
  a1 *= a1;                                                                          // :508 "need norm^2"
  UG_TRY(uggpu_dscal(ctx, level, level, UGGPU_ALL_VECTORS, c, 1 + a0 / a1));        // :511
  UG_TRY(uggpu_daxpy(ctx, level, level, UGGPU_ALL_VECTORS, b, -a0 / a1, t));        // :513
  return 0;
}

// the fused schedule is built on the Jacobi step; every other smoother class runs one kernel group per reference call, and so does
// the level optimisation of `transfer $L` (its scalars come back to the host between the calls)
static inline bool use_fused(const uggpu_lmgc_cfg *cfg) { return cfg->fused && cfg->smoother == UGGPU_SM_JAC && !cfg->level_opt; }

// t_ready: the temporary cfg->t of this level already holds damp * Diag(A)^-1 b (written by the restriction above)
// push_c: the caller interpolates this level's correction next -- on a partitioned level (multi-GPU, peer-memory ghost rows) the last
// kernel stores the interface rows of c into the neighbours' ghost rows.  The same holds for every producer -> consumer pair of the
// schedule (HaloPlan): the kernel that computes a vector pushes the rows the next kernel's ghost columns need, and the next kernel
// waits for its neighbours at its own head -- no exchange kernels in between.
static int lmgc_fused(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A, bool t_ready, TopFuse *tf, bool push_c = false, int t_buf = 1)
{
  if (level <= cfg->baselevel) return base_solve(ctx, cfg, level, c, b, A);
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  const int bs = L->bs;
  double *tA = get_vec(ctx, level, cfg->t), *tB = get_vec(ctx, level, UGGPU_VEC_TMP_B);
  double *cp = get_vec(ctx, level, c), *bp = get_vec(ctx, level, b);
  if (!tA || !tB || !cp || !bp) return UGGPU_DESC_MISMATCH;
  const Damp sd = mkdamp(cfg->smooth_damp, bs), cd = mkdamp(cfg->cycle_damp, bs), one = mkdamp(nullptr, 0);
  const bool top = tf && tf->level == level;
  bool c_zero = top && tf->c_zero;
  double *xp = nullptr;
  // x is only touched by the last smoothing step: an upload of it that is still in flight (uggpu_vec_upload_async) overlaps the cycle
  if (top) { xp = get_vec_lazy(ctx, level, tf->x); if (!xp) return UGGPU_DESC_MISMATCH; }
  double *cur = tA, *oth = tB;
  if (t_ready && t_buf == 2) { cur = tB; oth = tA; }      // the previous cycle's last step left the correction in the second temporary
  const bool part = halo_fused_available(ctx, level), cpart = halo_fused_available(ctx, level - 1);
  auto plan = [&](bool ready, double *push, int push_level) { return HaloPlan{ready, push, push_level}; };
  // Pairs of consecutive steps share one update of c (SF_CPREV): the first step of a pair leaves c alone, the second adds both corrections
  // in the reference's order.  One pair at the end of the pre-smoothing, one at the end of the post-smoothing; only where both steps run in
  // the stencil-rows / exception-rows kernels.
  SellMat *Mc = get_mat(ctx, level, A);
  const bool pairs = Mc && stx_handles(L, Mc) && !getenv("UGGPU_OVERLAP");

  if (cfg->nu1 > 0) {
    // the Jacobi start of the cycle's top level is a pure streaming kernel (measured: the flag byte of the comm form costs it 37 %): its
    // result goes to the neighbours by the stand-alone exchange in front of the first smoothing step
    if (!t_ready) UG_TRY(k_jac(ctx, level, A, cur, bp, sd, nullptr));
    for (int i = 0; i < cfg->nu1; i++) {
      const bool lastpre = i == cfg->nu1 - 1;
      const bool defer = pairs && cfg->nu1 >= 2 && i == cfg->nu1 - 2, second = pairs && cfg->nu1 >= 2 && lastpre;
      int flags = (defer ? 0 : (c_zero ? SF_CSET : SF_CADD) | (second ? SF_CPREV : 0)) | (!lastpre ? SF_TOUT : 0);
      const HaloPlan hp = plan(t_ready || i > 0, part ? (lastpre ? bp : oth) : nullptr, level);       // the last pre-smoothing step hands b to the restriction
      UG_TRY(k_smooth_step(ctx, level, A, flags, cur, bp, cp, oth, sd, nullptr, 0, &hp));
      if (!defer) c_zero = false;
      double *sw = cur; cur = oth; oth = sw;
    }
  }
  // restriction (+ first Jacobi correction and c = 0 of the coarse level)
  {
    const int lc = level - 1;
    double *bc = get_vec(ctx, lc, b), *cc = get_vec(ctx, lc, c), *tc = get_vec(ctx, lc, cfg->t);
    if (!bc || !cc || !tc) return UGGPU_DESC_MISMATCH;
    // not across the gather level of a multi-GPU hierarchy: there the coarse defect is complete only after the all-reduce
    const bool fuse = lc > cfg->baselevel && cfg->nu1 > 0 && !(ctx->comm && L->partitioned && !ctx->lev[lc].partitioned);
    {
      const HaloPlan hp = plan(cfg->nu1 > 0, cpart && lc > cfg->baselevel ? (fuse ? tc : bc) : nullptr, lc);
      UG_TRY(k_restrict(ctx, level, bc, bp, one, fuse, A, tc, cc, sd, &hp));
    }
    if (!fuse) UG_TRY(k_vec_op(ctx, lc, 0, VOP_SET, cc, nullptr, Damp{{0.0, 0.0, 0.0}}));   // dset(c,0) iter.cc:7873
    for (int g = 0; g < cfg->gamma; g++) UG_TRY(lmgc_fused(ctx, cfg, lc, c, b, A, fuse && g == 0, nullptr, g == cfg->gamma - 1));
    const HaloPlan hp = plan(true, part ? tA : nullptr, level);
    UG_TRY(k_interpolate(ctx, level, tA, cc, cd, &hp));
  }
  // c += t ; b -= A t ; first post-smoothing correction
  {
    const bool last = cfg->nu2 == 0;
    const bool defer = pairs && cfg->nu2 == 1;                  // with one post-smoothing step the pair is (this step, that step)
    int flags = (defer ? 0 : (c_zero ? SF_CSET : SF_CADD)) | (last ? 0 : SF_TOUT);
    bool nt = false;
    if (last && top) {
      flags |= SF_XADD | (tf->norm ? SF_NORM : 0); tf->done_x = true; tf->done_norm = tf->norm; UG_TRY(vec_wait(ctx, level, tf->x));
      if (tf->want_t && tf->norm && cfg->nu1 > 0) { flags |= SF_TOUT; tf->next_t = 2; nt = true; }
    }
    const HaloPlan hp = plan(true, part ? (last ? (nt ? tB : (push_c ? cp : nullptr)) : tB) : nullptr, level);
    UG_TRY(k_smooth_step(ctx, level, A, flags, tA, bp, cp, tB, sd, xp, 0, &hp));
    if (!defer) c_zero = false;
    cur = tB; oth = tA;
  }
  for (int i = 0; i < cfg->nu2; i++) {
    const bool last = i == cfg->nu2 - 1;
    const bool defer = pairs && cfg->nu2 >= 2 && i == cfg->nu2 - 2, second = pairs && last;
    int flags = (defer ? 0 : ((c_zero ? SF_CSET : SF_CADD) | (second ? SF_CPREV : 0))) | (last ? 0 : SF_TOUT);
    bool nt = false;
    if (last && top) {
      flags |= SF_XADD | (tf->norm ? SF_NORM : 0); tf->done_x = true; tf->done_norm = tf->norm; UG_TRY(vec_wait(ctx, level, tf->x));
      if (tf->want_t && tf->norm && cfg->nu1 > 0) { flags |= SF_TOUT; tf->next_t = oth == tA ? 1 : 2; nt = true; }
    }
    const HaloPlan hp = plan(true, part ? (last ? (nt ? oth : (push_c ? cp : nullptr)) : oth) : nullptr, level);
    UG_TRY(k_smooth_step(ctx, level, A, flags, cur, bp, cp, oth, sd, xp, 0, &hp));
    if (!defer) c_zero = false;
    double *sw = cur; cur = oth; oth = sw;
  }
  return 0;
}

static int lmgc_check(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b)
{
  if (!ctx || !cfg) return uggpu_fail(UGGPU_ERROR, "null argument");
  if (cfg->nu1 < 0 || cfg->nu2 < 0 || cfg->gamma < 1) return uggpu_fail(UGGPU_ERROR, "lmgc: bad nu1/nu2/gamma");
  for (int l = cfg->baselevel; l <= level; l++) {
    if (!get_level(ctx, l)) return UGGPU_ERROR;
    UG_TRY(ensure_vec(ctx, l, c));
    UG_TRY(ensure_vec(ctx, l, b));
    UG_TRY(ensure_vec(ctx, l, cfg->t));
    UG_TRY(ensure_vec(ctx, l, UGGPU_VEC_TMP_B));
  }
  return 0;
}

extern "C" int uggpu_lmgc(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A)
{
  UG_TRY(lmgc_check(ctx, cfg, level, c, b));
  if (use_fused(cfg)) UG_TRY(lmgc_fused(ctx, cfg, level, c, b, A, false, nullptr));
  else UG_TRY(lmgc_unfused(ctx, cfg, level, c, b, A));
  return check_device_error(ctx);
}

// ---- linear solver -------------------------------------------------------------------------------------------------------
extern "C" int uggpu_ls_defect(uggpu_ctx *ctx, int bl, int level, int x, int b, int A)
{
  (void)bl;   // ON_SURFACE loops start at FULLREFINELEVEL whatever fl is (matloop.ct:22)
  return uggpu_dmatmul_minus(ctx, bl, level, UGGPU_ON_SURFACE, b, A, x);
}

extern "C" int uggpu_ls_residuum(uggpu_ctx *ctx, int bl, int level, int b, uggpu_lresult *res)
{
  if (!res) return uggpu_fail(UGGPU_ERROR, "null result");
  double s[UGGPU_MAX_BS]; int bs;
  UG_TRY(reduce_loop(ctx, bl, level, UGGPU_ON_SURFACE, RED_NRM2, b, b, s, &bs));
  for (int i = 0; i < bs; i++) res->last_defect[i] = sqrt(s[i]);
  return 0;
}

extern "C" int uggpu_ls_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, int c,
                              int maxiter, const double *abslimit, const double *reduction, uggpu_lresult *res, double *history)
{
  if (!res || !abslimit || !reduction) return uggpu_fail(UGGPU_ERROR, "null argument");
  UG_TRY(lmgc_check(ctx, cfg, level, c, b));
  Level *L = get_level(ctx, level);
  const int bs = L->bs;
  for (int l = bl; l <= level; l++) if (!get_vec_lazy(ctx, l, x)) return UGGPU_DESC_MISMATCH;
  double reach[UGGPU_MAX_BS];
  res->error_code = 0; res->converged = 0; res->number_of_linear_iterations = 0;
  for (int i = 0; i < bs; i++) {
    res->first_defect[i] = res->last_defect[i];
    reach[i] = res->last_defect[i] * reduction[i];
    if (reach[i] == 0.0) reach[i] = reduction[i];              // ls.cc:663-667
  }
  if (sc_cmp(res->last_defect, abslimit, bs)) { res->converged = 1; return 0; }
  const Damp none = mkdamp(nullptr, 0);
  const bool surface_is_top = ctx->fullrefinelevel >= level;   // no lower-level terms in the ON_SURFACE norm
  int next_t = 0;                                              // the previous iteration's last kernel left the next Jacobi correction there (TopFuse)
  for (int it = 0; it < maxiter; it++) {
    bool done_x = false, done_norm = false;
    if (use_fused(cfg) && level > cfg->baselevel) {
      TopFuse tf;
      tf.level = level; tf.x = x; tf.c_zero = true; tf.norm = surface_is_top;
      tf.want_t = it + 1 < maxiter && !getenv("UGGPU_NO_NEXT_T");
      UG_TRY(lmgc_fused(ctx, cfg, level, c, b, A, next_t != 0, &tf, false, next_t ? next_t : 1));
      next_t = tf.next_t;
      done_x = tf.done_x; done_norm = tf.done_norm;
    } else {
      UG_TRY(uggpu_dset(ctx, level, level, UGGPU_ALL_VECTORS, c, 0.0));     // ls.cc:695
      if (use_fused(cfg)) UG_TRY(lmgc_fused(ctx, cfg, level, c, b, A, false, nullptr));
      else UG_TRY(lmgc_unfused(ctx, cfg, level, c, b, A));
    }
    // LSUpdate (ls.cc:869): x += c on levels bl..level
    for (int l = bl; l <= level; l++) {
      if (l == level && done_x) continue;
      UG_TRY(k_vec_op(ctx, l, 0, VOP_ADD, get_vec(ctx, l, x), get_vec(ctx, l, c), none));
    }
    if (done_norm) {
      UG_TRY(fetch_results(ctx, 1));
      for (int i = 0; i < bs; i++) res->last_defect[i] = sqrt(ctx->hres[i]);
    } else {
      UG_TRY(uggpu_ls_residuum(ctx, bl, level, b, res));
    }
    if (history) for (int i = 0; i < bs; i++) history[it * bs + i] = res->last_defect[i];
    res->number_of_linear_iterations = it + 1;
    if (sc_cmp(res->last_defect, abslimit, bs) || sc_cmp(res->last_defect, reach, bs)) { res->converged = 1; break; }
  }
  int rc = check_device_error(ctx);
  if (rc) res->error_code = rc;
  return rc;
}

// ---- Krylov accelerators around the cycle ------------------------------------------------------------------
// The reference's classes `cg` (LinearSolver ls.cc:637 with CGPrepare :976 / CGUpdate :989-1027 / CGClose :1159) and `bcgs`
// (BCGSSolver ls.cc:1864-2062) with Iter = the cycle, every operation a device kernel of this library in the reference's
// order; only the scalars (lambda, rho, alpha, omega, the defect norms) travel to the host.  Vectors agree with the reference
// to rounding, not bit for bit: the scalars come from parallel sums.
static int run_cycle(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int level, int c, int b, int A)
{
  if (use_fused(cfg)) return lmgc_fused(ctx, cfg, level, c, b, A, false, nullptr);
  return lmgc_unfused(ctx, cfg, level, c, b, A);
}

static int krylov_begin(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int c, const int *work, int nwork,
                        const double *abslimit, const double *reduction, uggpu_lresult *res, double *reach, int *bs_out)
{
  if (!res || !abslimit || !reduction) return uggpu_fail(UGGPU_ERROR, "null argument");
  UG_TRY(lmgc_check(ctx, cfg, level, c, b));
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  for (int l = bl; l <= level; l++) {
    if (!get_vec(ctx, l, x)) return UGGPU_DESC_MISMATCH;
    for (int k = 0; k < nwork; k++) UG_TRY(ensure_vec(ctx, l, work[k]));
  }
  const int bs = L->bs;
  res->error_code = 0; res->converged = 0; res->number_of_linear_iterations = 0;
  for (int i = 0; i < bs; i++) {
    res->first_defect[i] = res->last_defect[i];
    reach[i] = res->last_defect[i] * reduction[i];
    if (reach[i] == 0.0) reach[i] = reduction[i];              // sc_mul_check, ls.cc:663-667
  }
  *bs_out = bs;
  return 0;
}

// Fused BLAS-1 chains (blas1.cu chain_loop) in cg / bcgs: on unless UGGPU_NO_KRYLOV_FUSION is set (A/B: identical results) or the solve
// starts above FULLREFINELEVEL (the ON_SURFACE reductions then cover levels the ALL_VECTORS operations do not)
static bool krylov_fusion(uggpu_ctx *ctx, int bl) { return !getenv("UGGPU_NO_KRYLOV_FUSION") && bl <= ctx->fullrefinelevel; }

extern "C" int uggpu_ddotw(uggpu_ctx *ctx, int fl, int tl, int mode, int x, int y, const double *w, double *a)
{
  double s[UGGPU_MAX_BS]; int bs;
  if (!w || !a) return uggpu_fail(UGGPU_ERROR, "null argument");
  UG_TRY(reduce_loop(ctx, fl, tl, mode, RED_DOT, x, y, s, &bs));
  *a = 0.0;
  for (int i = 0; i < bs; i++) *a += w[i] * s[i];              // T_POST of ddotw, ugblas.cc:3044
  return 0;
}

extern "C" int uggpu_cg_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, int c, int p, int t,
                              int maxiter, const double *abslimit, const double *reduction, uggpu_lresult *res, double *history)
{
  double reach[UGGPU_MAX_BS]; int bs;
  const int work[2] = {p, t};
  UG_TRY(krylov_begin(ctx, cfg, bl, level, x, b, c, work, 2, abslimit, reduction, res, reach, &bs));
  const int ALL = UGGPU_ALL_VECTORS, SURF = UGGPU_ON_SURFACE;
  UG_TRY(uggpu_dset(ctx, bl, level, ALL, p, 0.0));                                  // CGPrepare
  double rho = 1.0, lambda = 0.0;
  const bool fuse = krylov_fusion(ctx, bl);
  if (sc_cmp(res->last_defect, abslimit, bs)) { res->converged = 1; return 0; }
  for (int it = 0; it < maxiter; it++) {
    UG_TRY(uggpu_dset(ctx, level, level, ALL, c, 0.0));                             // ls.cc:695
    UG_TRY(run_cycle(ctx, cfg, level, c, b, A));
    UG_TRY(uggpu_dmatmul(ctx, bl, level, ALL, t, A, c));                            // CGUpdate :1003
    if (fuse) {
      // the BLAS-1 calls of CGUpdate in three passes instead of seven (same operations per entry, same partial sums: bit-identical)
      double s3[UGGPU_MAX_BS];
      const int v1[3] = {b, t, c};
      UG_TRY(chain_loop(ctx, bl, level, CH_ADD_DOT, v1, 0.0, 0.0, s3, nullptr));      // dadd b t; ddot c b
      lambda = 0.0; for (int i = 0; i < bs; i++) lambda += s3[i];
      const int v2[2] = {p, c};
      UG_TRY(chain_loop(ctx, bl, level, CH_SCAL_ADD, v2, lambda / rho, 0.0, nullptr, nullptr));   // dscal p; dadd p c
      rho = lambda;
      UG_TRY(uggpu_dmatmul(ctx, bl, level, ALL, t, A, p));
      UG_TRY(uggpu_ddot(ctx, bl, level, SURF, t, p, &lambda));
      if (lambda == 0.0) { res->error_code = UGGPU_ERROR; return uggpu_fail(UGGPU_ERROR, "cg: (Ap,p) = 0 in iteration %d", it); }   // :1017
      const int v3[4] = {x, p, b, t};
      UG_TRY(chain_loop(ctx, bl, level, CH_AXPY2_NRM, v3, rho / lambda, -rho / lambda, s3, nullptr));   // daxpy x p; daxpy b t; LinearResiduum
      for (int i = 0; i < bs; i++) res->last_defect[i] = sqrt(s3[i]);
    } else {
    UG_TRY(uggpu_dadd(ctx, bl, level, ALL, b, t));
    UG_TRY(uggpu_ddot(ctx, bl, level, SURF, c, b, &lambda));
    UG_TRY(uggpu_dscal(ctx, bl, level, ALL, p, lambda / rho));
    rho = lambda;
    UG_TRY(uggpu_dadd(ctx, bl, level, ALL, p, c));
    UG_TRY(uggpu_dmatmul(ctx, bl, level, ALL, t, A, p));
    UG_TRY(uggpu_ddot(ctx, bl, level, SURF, t, p, &lambda));
    if (lambda == 0.0) { res->error_code = UGGPU_ERROR; return uggpu_fail(UGGPU_ERROR, "cg: (Ap,p) = 0 in iteration %d", it); }   // :1017
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, x, rho / lambda, p));
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, b, -rho / lambda, t));
    UG_TRY(uggpu_ls_residuum(ctx, bl, level, b, res));
    }
    if (history) for (int i = 0; i < bs; i++) history[it * bs + i] = res->last_defect[i];
    res->number_of_linear_iterations = it + 1;
    if (sc_cmp(res->last_defect, abslimit, bs) || sc_cmp(res->last_defect, reach, bs)) { res->converged = 1; break; }
  }
  int rc = check_device_error(ctx);
  if (rc) res->error_code = rc;
  return rc;
}

extern "C" int uggpu_bcgs_solve(uggpu_ctx *ctx, const uggpu_lmgc_cfg *cfg, int bl, int level, int x, int b, int A, const int *work,
                                const double *weight, int restart_every, int maxiter, const double *abslimit, const double *reduction,
                                uggpu_lresult *res, double *history)
{
  if (!work || !weight) return uggpu_fail(UGGPU_ERROR, "null argument");
  double reach[UGGPU_MAX_BS], old[UGGPU_MAX_BS] = {-1.0, -1.0, -1.0}; int bs;
  const int r = work[0], p = work[1], v = work[2], s = work[3], t = work[4], q = work[5];
  UG_TRY(krylov_begin(ctx, cfg, bl, level, x, b, q, work, 6, abslimit, reduction, res, reach, &bs));
  const int ALL = UGGPU_ALL_VECTORS, SURF = UGGPU_ON_SURFACE;
  double w2[UGGPU_MAX_BS];
  for (int i = 0; i < UGGPU_MAX_BS; i++) w2[i] = weight[i] * weight[i];             // BCGSInit :1757
  double alpha = 0.0, rho_new = 0.0, beta = 0.0, tt = 0.0, rho = 0.0, omega = 0.0;
  const bool fuse = krylov_fusion(ctx, bl);
  int restart = 1, eq_count = 0;
  if (sc_cmp(res->last_defect, abslimit, bs)) res->converged = 1;
  for (int i = 0; i < maxiter; i++) {
    if (res->converged) break;
    if ((restart_every > 0 && i % restart_every == 0) || restart) {
      UG_TRY(uggpu_dset(ctx, bl, level, ALL, p, 0.0));
      UG_TRY(uggpu_dset(ctx, bl, level, ALL, v, 0.0));
      UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, r, b));
      alpha = rho = omega = 1.0;
      restart = 0;
    }
    UG_TRY(uggpu_ddotw(ctx, bl, level, SURF, b, r, w2, &rho_new));
    if (rho != 0.0 && omega != 0.0) beta = rho_new * alpha / rho / omega;
    if (fuse) {
      const int v5[5] = {p, b, v, q, s};
      UG_TRY(chain_loop(ctx, bl, level, CH_BCGS_P, v5, beta, -beta * omega, nullptr, nullptr));    // the five calls below in one pass
    } else {
    UG_TRY(uggpu_dscal(ctx, bl, level, ALL, p, beta));
    UG_TRY(uggpu_dadd(ctx, bl, level, ALL, p, b));
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, p, -beta * omega, v));
    UG_TRY(uggpu_dset(ctx, bl, level, ALL, q, 0.0));
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, s, p));
    }
    UG_TRY(run_cycle(ctx, cfg, level, q, p, A));                                    // Iter(q, p) :1944
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, p, s));
    UG_TRY(uggpu_dmatmul(ctx, bl, level, SURF, v, A, q));
    UG_TRY(uggpu_ddotw(ctx, bl, level, SURF, v, r, w2, &alpha));
    if (alpha != 0.0) alpha = rho_new / alpha;
    res->number_of_linear_iterations++;
    if (fuse) {
      double s3[UGGPU_MAX_BS];
      const int v5[5] = {x, q, s, b, v};
      UG_TRY(chain_loop(ctx, bl, level, CH_BCGS_S, v5, alpha, -alpha, s3, nullptr));                // daxpy x q; dcopy s b; daxpy s v; LinearResiduum(s)
      for (int k = 0; k < bs; k++) res->last_defect[k] = sqrt(s3[k]);
    } else {
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, x, alpha, q));
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, s, b));
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, s, -alpha, v));
    UG_TRY(uggpu_ls_residuum(ctx, bl, level, s, res));
    }
    if (sc_cmp(res->last_defect, abslimit, bs) || sc_cmp(res->last_defect, reach, bs)) {
      UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, b, s));
      res->converged = 1;
      if (history) for (int k = 0; k < bs; k++) history[i * bs + k] = res->last_defect[k];
      break;
    }
    UG_TRY(uggpu_dset(ctx, bl, level, ALL, q, 0.0));
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, t, s));
    UG_TRY(run_cycle(ctx, cfg, level, q, s, A));                                    // Iter(q, s) :1991
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, s, t));
    UG_TRY(uggpu_dmatmul(ctx, bl, level, SURF, t, A, q));
    UG_TRY(uggpu_ddotw(ctx, bl, level, SURF, t, t, w2, &tt));
    UG_TRY(uggpu_ddotw(ctx, bl, level, SURF, s, t, w2, &omega));
    if (tt != 0.0) omega /= tt;
    rho = rho_new;
    if (fuse) {
      double s3[UGGPU_MAX_BS];
      const int v5[5] = {x, q, b, s, t};
      UG_TRY(chain_loop(ctx, bl, level, CH_BCGS_S, v5, omega, -omega, s3, nullptr));                // daxpy x q; dcopy b s; daxpy b t; LinearResiduum(b)
      for (int k = 0; k < bs; k++) res->last_defect[k] = sqrt(s3[k]);
    } else {
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, x, omega, q));
    UG_TRY(uggpu_dcopy(ctx, bl, level, ALL, b, s));
    UG_TRY(uggpu_daxpy(ctx, bl, level, ALL, b, -omega, t));
    UG_TRY(uggpu_ls_residuum(ctx, bl, level, b, res));
    }
    if (history) for (int k = 0; k < bs; k++) history[i * bs + k] = res->last_defect[k];
    res->number_of_linear_iterations++;
    if (sc_cmp(res->last_defect, abslimit, bs) || sc_cmp(res->last_defect, reach, bs)) { res->converged = 1; break; }
    int eq = 1;                                                                     // sc_eq(last, old, 1e-4), npscan.cc:1095
    for (int k = 0; k < bs; k++) {
      const double a = res->last_defect[k], o = old[k];
      if (a < 0.0 || o < 0.0 || fabs(a - o) > 1e-4 * sqrt(a * o)) eq = 0;
    }
    eq_count = eq ? eq_count + 1 : 0;
    for (int k = 0; k < bs; k++) old[k] = res->last_defect[k];
    if (eq_count > 4) { res->converged = 0; break; }
  }
  int rc = check_device_error(ctx);
  if (rc) res->error_code = rc;
  return rc;
}
