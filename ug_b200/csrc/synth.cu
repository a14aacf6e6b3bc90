// synth.cu -- synthetic multigrid hierarchies generated directly in device memory (bench / large-scale test input).
//
// There is no reference analogue: UG's grid manager needs ~2.5 kB of host memory per unknown, so
// hierarchies with >= 10^8 unknowns cannot be built by the reference and flattened.  The generator emits exactly
// the data model `gpuls` PreProcess extracts from UG (canonical layout of SURVEY.md 8a'): per level a BSR matrix
// with the diagonal first, per-row VCLASS/VNCLASS/NEW_DEFECT/FINE_GRID_DOF/VECSKIP flags as UG sets them on a
// uniformly refined multigrid (observed in tests/golden/*: levels below the top have class 3 / nclass 3 /
// NEW_DEFECT only, the top level nclass 0 and both bits, FULLREFINELEVEL = top, Dirichlet rows are identity
// rows with their connections kept at value 0 and all VECSKIP bits set), the standard prolongation P
// (corner nodes weight 1, other nodes the P1/Q1 shape-function values of the father element, zeros dropped) and
// the restriction R with entries in ascending (global) fine-row order.
//
// Meshes: structured nx x ny (x nz) cubic cells of size h0 = 1/nx on the box [0,1] x [0,ny/nx] (x [0,nz/nx]), rows
// numbered lexicographically (x fastest).
//   P1 simplices: every cell is cut into the 2 (6) Kuhn simplices around the diagonal 0-2 (0-6) (the coarse grids of
//     SURVEY.md Appendix A); finer levels are again Kuhn triangulations (Bey refinement), so every interior row has the
//     7 (15) connections along the directions {0,1}^d \ 0 and their negatives.  In 2D this is exactly the mesh UG's
//     own refinement produces (tests/test_synth.py matches it entry by entry against the golden dump); in 3D UG's
//     regular rule picks varying octahedron diagonals and ends with 14.6 connections per row on average instead of 15.
//     P1 Laplace on a Kuhn mesh: diagonal 2d h^(d-2), axis neighbours -h^(d-2), diagonal neighbours 0 (stored).
//
//   Q1 cubes (scalar Poisson or 3x3 linear elasticity, E = 1, nu = 0.3): rows assembled on the device from the 8x8
//     (24x24) element matrix integrated on the host with 2x2x2 Gauss points; 27 block connections per interior row;
//     UG's regular hexahedron rule produces exactly this mesh (tests/test_synth.py matches the golden dump).
//
// Multi-GPU: with a Px x Py x Pz rank array (uggpu_synth_hierarchy_part) every rank generates only the rows it owns
// plus ghost columns (part.h); levels with at most `replicate_below` rows are generated completely on every rank.
// The arithmetic per row is the same as on one GPU, so partitioned and unpartitioned solves agree bit for bit.
#include "uggpu_internal.h"
#include "part.h"

#include <cmath>
#include <cstring>
#include <vector>

struct SynthParams {
  int kind;         // UGGPU_SYNTH_*
  int dim;
  int bs;           // components per node
  double hpow;      // h^(dim-2): scale of the stiffness entries
  double hvol;      // h^dim
  double hhalf;     // h / 2 (edge midpoints, UGGPU_SYNTH_P1_VARCOEF)
};

__device__ __forceinline__ bool in_grid(const int (&nn)[3], const int (&x)[3]) { return x[0] >= 0 && x[0] < nn[0] && x[1] >= 0 && x[1] < nn[1] && x[2] >= 0 && x[2] < nn[2]; }
__device__ __forceinline__ bool on_boundary(const PartGrid &g, const int (&x)[3])
{
  for (int d = 0; d < g.dim; d++) if (x[d] == 0 || x[d] == g.nn[d] - 1) return true;
  return false;
}

// Kuhn directions in canonical entry order (after the diagonal): +d, -d for d = e1, e2, e3, e1+e2, e2+e3, e1+e3, e1+e2+e3
__constant__ int c_kuhn3[7][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
__constant__ int c_kuhn2[3][3] = {{1, 0, 0}, {0, 1, 0}, {1, 1, 0}};
// Q1 element matrix of the unit cube (h = 1), local node a = ox + 2 oy + 4 oz, row-major (8 bs) x (8 bs); the matrix of
// a cube of size h is h times this one (gradients ~ 1/h, volume ~ h^3)
__constant__ double c_ke[24 * 24];

#define SYNTH_MAXROW 27

struct RowGen {
  int len;
  int32_t cols[SYNTH_MAXROW];
  int8_t code[SYNTH_MAXROW];   // hex matrix rows: neighbour offset (qx+1) + 3(qy+1) + 9(qz+1)
  double w[SYNTH_MAXROW];      // scalar value of the entry (simplex matrix rows, P and R rows)
};

__device__ __forceinline__ bool is_simplex(int kind) { return kind == UGGPU_SYNTH_P1_SIMPLEX || kind == UGGPU_SYNTH_P1_VARCOEF; }

// UGGPU_SYNTH_P1_VARCOEF: diffusion coefficient on the edge x--y, a smooth function of the edge midpoint in physical coordinates
// (1 <= kappa < 2.2; symmetric in x and y; the same doubles on every rank and level-independent as a field): -div(kappa grad u)
// in the two-point-flux form on the axis edges of the Kuhn mesh.  No two rows of a slice share their values.
__device__ __forceinline__ double edge_kappa(const SynthParams &sp, const int (&x)[3], const int (&y)[3])
{
  const double u = (double)(x[0] + y[0]) * sp.hhalf, v = (double)(x[1] + y[1]) * sp.hhalf, w = (double)(x[2] + y[2]) * sp.hhalf;
  double k = 1.0 + 0.9 * (u * (1.0 - u));
  k = k + 0.7 * (v * v);
  k = k + 0.5 * (u * w);
  return k;
}

// one row of the P1 simplex matrix
__device__ void simplex_row(const SynthParams &sp, const PartGrid &g, int r, RowGen &rg)
{
  int x[3];
  if (!part_row_coords(g, r, x)) { rg.len = 0; return; }      // dummy row of the padded numbering (part.h): empty
  const bool bnd = on_boundary(g, x);
  const bool var = sp.kind == UGGPU_SYNTH_P1_VARCOEF;
  int len = 0;
  rg.cols[len] = r; rg.w[len] = bnd ? 1.0 : 2.0 * sp.dim * sp.hpow; len++;
  const int nd = sp.dim == 3 ? 7 : 3;
  double dsum = 0.0;
  for (int k = 0; k < nd; k++) {
    const int *d = sp.dim == 3 ? c_kuhn3[k] : c_kuhn2[k];
    const bool axis = (d[0] + d[1] + d[2]) == 1;
    for (int sgn = 1; sgn >= -1; sgn -= 2) {
      int y[3] = {x[0] + sgn * d[0], x[1] + sgn * d[1], x[2] + sgn * d[2]};
      if (!in_grid(g.nn, y)) continue;
      rg.cols[len] = part_local_index(g, y);
      double w = (bnd || !axis) ? 0.0 : -sp.hpow;
      if (var && w != 0.0) { const double kw = edge_kappa(sp, x, y) * sp.hpow; w = -kw; dsum += kw; }
      rg.w[len] = w;
      len++;
    }
  }
  if (var && !bnd) rg.w[0] = dsum;
  rg.len = len;
}

// pattern of one row of a Q1 matrix: diagonal, then the up to 26 lattice neighbours in lexicographic order
__device__ void hex_row(const SynthParams &sp, const PartGrid &g, int r, RowGen &rg)
{
  int x[3];
  if (!part_row_coords(g, r, x)) { rg.len = 0; return; }
  int len = 0;
  rg.cols[len] = r; rg.code[len] = 13; len++;
  for (int qz = -1; qz <= 1; qz++)
    for (int qy = -1; qy <= 1; qy++)
      for (int qx = -1; qx <= 1; qx++) {
        if (!(qx | qy | qz)) continue;
        int y[3] = {x[0] + qx, x[1] + qy, x[2] + qz};
        if (!in_grid(g.nn, y)) continue;
        rg.cols[len] = part_local_index(g, y);
        rg.code[len] = (int8_t)((qx + 1) + 3 * (qy + 1) + 9 * (qz + 1));
        len++;
      }
  rg.len = len;
}

// bs x bs block of the Q1 matrix between node x and its neighbour `code`: sum of the element matrices of the (up to 8)
// cells containing both, in fixed (oz, oy, ox) order; Dirichlet rows are identity rows
__device__ void hex_block(const SynthParams &sp, const PartGrid &g, const int (&x)[3], int code, double *blk)
{
  const int bs = sp.bs, ld = 8 * bs;
  for (int k = 0; k < bs * bs; k++) blk[k] = 0.0;
  if (on_boundary(g, x)) {
    if (code == 13) for (int i = 0; i < bs; i++) blk[i * bs + i] = 1.0;
    return;
  }
  const int q[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
  for (int oz = 0; oz <= 1; oz++)
    for (int oy = 0; oy <= 1; oy++)
      for (int ox = 0; ox <= 1; ox++) {
        const int o[3] = {ox, oy, oz};
        bool ok = true;
        int b[3];
        for (int d = 0; d < 3; d++) {
          int c = x[d] - o[d];
          if (c < 0 || c > g.nn[d] - 2) ok = false;
          b[d] = o[d] + q[d];
          if (b[d] < 0 || b[d] > 1) ok = false;
        }
        if (!ok) continue;
        const int la = ox + 2 * oy + 4 * oz, lb = b[0] + 2 * b[1] + 4 * b[2];
        for (int i = 0; i < bs; i++)
          for (int j = 0; j < bs; j++) blk[i * bs + j] += c_ke[(la * bs + i) * ld + lb * bs + j];
      }
  for (int k = 0; k < bs * bs; k++) blk[k] *= sp.hpow;
}

// P row of fine node r.  Simplices: corner node -> (father node, 1), else the midpoint of the Kuhn edge along its parity
// vector.  Cubes: trilinear weights 2^-|p| of the 2^|p| coarse nodes around it (lexicographic).
__device__ void p_row(const SynthParams &sp, const PartGrid &g, const PartGrid &gc, int r, RowGen &rg)
{
  int x[3];
  if (!part_row_coords(g, r, x)) { rg.len = 0; return; }
  const int p[3] = {x[0] & 1, x[1] & 1, x[2] & 1};
  const int np = p[0] + p[1] + p[2];
  if (np == 0) {
    int X[3] = {x[0] >> 1, x[1] >> 1, x[2] >> 1};
    rg.cols[0] = part_local_index(gc, X); rg.w[0] = 1.0; rg.len = 1;
    return;
  }
  if (is_simplex(sp.kind)) {
    int a[3] = {(x[0] - p[0]) >> 1, (x[1] - p[1]) >> 1, (x[2] - p[2]) >> 1};
    int b[3] = {(x[0] + p[0]) >> 1, (x[1] + p[1]) >> 1, (x[2] + p[2]) >> 1};
    rg.cols[0] = part_local_index(gc, a); rg.w[0] = 0.5;
    rg.cols[1] = part_local_index(gc, b); rg.w[1] = 0.5;
    rg.len = 2;
    return;
  }
  const double w = np == 1 ? 0.5 : (np == 2 ? 0.25 : 0.125);
  int len = 0;
  for (int ez = 0; ez <= p[2]; ez++)
    for (int ey = 0; ey <= p[1]; ey++)
      for (int ex = 0; ex <= p[0]; ex++) {
        int X[3] = {(x[0] - p[0]) / 2 + ex, (x[1] - p[1]) / 2 + ey, (x[2] - p[2]) / 2 + ez};
        rg.cols[len] = part_local_index(gc, X); rg.w[len] = w; len++;
      }
  rg.len = len;
}

// R row of coarse row R: fine nodes 2X+q in ascending global fine index.  Simplices: q in {all components >= 0} u
// {all <= 0} (the Kuhn edges at X), weight 1/2; cubes: all q in {-1,0,1}^3, weight 2^-|q|.
// On the gather level of a multi-GPU hierarchy (fine partitioned, coarse complete) only the rows of the nodes this rank
// would own are filled: the all-reduce that follows the restriction kernel adds the disjoint parts.
__device__ void r_row(const SynthParams &sp, const PartGrid &g, const PartGrid &gc, int R, RowGen &rg)
{
  int X[3];
  rg.len = 0;
  if (!part_row_coords(gc, R, X)) return;
  if (!g.replicated && gc.replicated && !box_has(gc.own, X)) return;
  int len = 0;
  const int z0 = sp.dim == 3 ? -1 : 0, z1 = sp.dim == 3 ? 1 : 0;
  for (int qz = z0; qz <= z1; qz++)
    for (int qy = -1; qy <= 1; qy++)
      for (int qx = -1; qx <= 1; qx++) {
        if (is_simplex(sp.kind)) {
          const bool nonneg = qx >= 0 && qy >= 0 && qz >= 0, nonpos = qx <= 0 && qy <= 0 && qz <= 0;
          if (!nonneg && !nonpos) continue;
        }
        int y[3] = {2 * X[0] + qx, 2 * X[1] + qy, 2 * X[2] + qz};
        if (!in_grid(g.nn, y)) continue;
        const int nq = (qx != 0) + (qy != 0) + (qz != 0);
        rg.cols[len] = part_local_index(g, y);
        rg.w[len] = is_simplex(sp.kind) ? (nq ? 0.5 : 1.0) : (nq == 0 ? 1.0 : (nq == 1 ? 0.5 : (nq == 2 ? 0.25 : 0.125)));
        len++;
      }
  rg.len = len;
}

enum { GEN_A = 0, GEN_P = 1, GEN_R = 2 };

template <int WHICH>
__device__ __forceinline__ void gen_row(const SynthParams &sp, const PartGrid &g, const PartGrid &gc, int r, RowGen &rg)
{
  if (WHICH == GEN_A) { if (is_simplex(sp.kind)) simplex_row(sp, g, r, rg); else hex_row(sp, g, r, rg); }
  else if (WHICH == GEN_P) p_row(sp, g, gc, r, rg);
  else r_row(sp, g, gc, r, rg);
}

template <int WHICH>
__global__ void k_synth_len(SynthParams sp, const PartGrid *__restrict__ g, const PartGrid *__restrict__ gc, int n, uint16_t *__restrict__ rowlen,
                            int *__restrict__ width, unsigned long long *nnz)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  RowGen rg;
  rg.len = 0;
  if (r < n) { gen_row<WHICH>(sp, *g, *gc, r, rg); rowlen[r] = (uint16_t)rg.len; }
  int w = rg.len, s = rg.len;
  for (int o = 16; o > 0; o >>= 1) { w = max(w, __shfl_xor_sync(0xffffffffu, w, o)); s += __shfl_xor_sync(0xffffffffu, s, o); }
  if ((threadIdx.x & 31) == 0 && (r >> 5) < (n + 31) / 32) { width[r >> 5] = w; atomicAdd(nnz, (unsigned long long)s); }
}

template <int WHICH>
__global__ void k_synth_fill(SynthParams sp, const PartGrid *__restrict__ g, const PartGrid *__restrict__ gc, int n, const int64_t *__restrict__ slice_ptr,
                             int32_t *__restrict__ col, double *__restrict__ val)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  int s = r >> 5, lane = r & 31;
  if (s >= (n + 31) / 32) return;
  RowGen rg;
  rg.len = 0;
  if (r < n) gen_row<WHICH>(sp, *g, *gc, r, rg);
  const bool blocks = WHICH == GEN_A && !is_simplex(sp.kind);
  const int bb = blocks ? sp.bs * sp.bs : 1;
  int x[3] = {0, 0, 0};
  if (blocks && r < n) (void)part_row_coords(*g, r, x);
  const int64_t spt = slice_ptr[s];
  const int w = (int)((slice_ptr[s + 1] - spt) >> 5);
  const int padcol = r < n ? r : 0;
  for (int j = 0; j < w; j++) {
    col[spt + (int64_t)j * 32 + lane] = j < rg.len ? rg.cols[j] : padcol;
    double blk[UGGPU_MAX_BS * UGGPU_MAX_BS];
    if (j < rg.len) { if (blocks) hex_block(sp, *g, x, rg.code[j], blk); else blk[0] = rg.w[j]; }
    else for (int k = 0; k < bb; k++) blk[k] = 0.0;
    for (int k = 0; k < bb; k++) val[(spt + (int64_t)j * 32) * bb + (int64_t)k * 32 + lane] = blk[k];
  }
}

__global__ void k_synth_flags(const PartGrid *__restrict__ g, int n, int bs, bool top, uint8_t *vclass, uint8_t *vnclass, uint8_t *ctl, uint32_t *skip)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int x[3];
  if (!part_row_coords(*g, r, x)) { vclass[r] = 0; vnclass[r] = 0; ctl[r] = 0; skip[r] = (1u << bs) - 1u; return; }      // dummy row: inert
  vclass[r] = 3;
  vnclass[r] = top ? 0 : 3;
  ctl[r] = top ? (UGGPU_CTL_NEW_DEFECT | UGGPU_CTL_FINE_GRID_DOF) : UGGPU_CTL_NEW_DEFECT;
  skip[r] = on_boundary(*g, x) ? ((1u << bs) - 1u) : 0u;
}

// load vector: scalar problems f = 1; elasticity: unit gravity on the last component (as oracle/ug_driver.cc assembles it)
__global__ void k_synth_rhs(SynthParams sp, const PartGrid *__restrict__ g, int n, double *b)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int x[3];
  const bool bnd = !part_row_coords(*g, r, x) || on_boundary(*g, x);      // dummy rows: 0
  for (int i = 0; i < sp.bs; i++) {
    double v = 0.0;
    if (!bnd) { if (sp.bs == 1) v = sp.hvol; else if (i == sp.bs - 1) v = -sp.hvol; }
    b[(size_t)r * sp.bs + i] = v;
  }
}

// global lexicographic id of every owned row (tests: assemble a global vector from the ranks' parts)
__global__ void k_synth_gids(const PartGrid *__restrict__ g, int n, int64_t *ids)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  int x[3];
  if (!part_row_coords(*g, r, x)) { ids[r] = -1; return; }      // dummy row of the padded numbering
  ids[r] = (int64_t)x[0] + (int64_t)g->nn[0] * ((int64_t)x[1] + (int64_t)g->nn[1] * x[2]);
}

// local indices of the rows to send, neighbour by neighbour (box order = the receiver's ghost order)
__global__ void k_synth_sendidx(const PartGrid *__restrict__ g, int total, int32_t *idx)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int k = 0;
  while (i >= g->nb_send_off[k + 1]) k++;
  int x[3];
  box_unlex(g->nb_send[k], i - g->nb_send_off[k], x);
  idx[i] = part_own_lex(*g, x);
}

template <int WHICH>
static int synth_sell(uggpu_ctx *ctx, const SynthParams &sp, const PartGrid *d_g, const PartGrid *d_gc, int n, SellMat *out)
{
  cudaStream_t st = ctx->stream;
  SellMat m;
  m.n = n; m.bb = (WHICH == GEN_A) ? sp.bs * sp.bs : 1;
  size_t nsl = (size_t)(n + 31) / 32;
  int *d_width = nullptr;
  unsigned long long *d_nnz = nullptr, h_nnz = 0;
  UG_TRY(dalloc(ctx, &m.rowlen, (size_t)n));
  UG_TRY(dalloc(ctx, &m.slice_ptr, nsl + 1));
  UG_TRY(dalloc(ctx, &d_width, nsl));
  UG_TRY(dalloc(ctx, &d_nnz, 1));
  CUDA_TRY(cudaMemsetAsync(d_nnz, 0, sizeof(unsigned long long), st));
  std::vector<int> width(nsl);
  std::vector<int64_t> sp_host;
  int blocks = (int)((nsl * 32 + 255) / 256);
  if (blocks > 0) {
    k_synth_len<WHICH><<<blocks, 256, 0, st>>>(sp, d_g, d_gc, n, m.rowlen, d_width, d_nnz);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaMemcpyAsync(width.data(), d_width, nsl * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&h_nnz, d_nnz, sizeof h_nnz, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  m.nnz = (int64_t)h_nnz;
  sell_layout(width, sp_host, &m.maxlen, &m.fixed_w);
  m.padded = sp_host[nsl];
  UG_TRY(dfree(ctx, d_width, nsl));
  UG_TRY(dfree(ctx, d_nnz, 1));
  CUDA_TRY(cudaMemcpyAsync(m.slice_ptr, sp_host.data(), (nsl + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  UG_TRY(dalloc(ctx, &m.col, (size_t)m.padded));
  UG_TRY(dalloc(ctx, &m.val, (size_t)m.padded * m.bb));
  if (blocks > 0) {
    k_synth_fill<WHICH><<<blocks, 256, 0, st>>>(sp, d_g, d_gc, n, m.slice_ptr, m.col, m.val);
    KCHECK(ctx);
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  m.col_ptr = m.slice_ptr; m.col_len = m.padded; m.col_words = m.nnz;
  UG_TRY(sell_compress_cols(ctx, &m));
  *out = m;
  return 0;
}

// Q1 element matrix of the unit cube with 2x2x2 Gauss points, exactly as oracle/ug_driver.cc assembles it:
// scalar: grad N_a . grad N_b;  elasticity: lam g_a[i] g_b[j] + mu g_a[j] g_b[i] + delta_ij mu g_a . g_b  (E = 1, nu = 0.3)
static void hex_element_matrix(int bs, double *ke /* (8 bs)^2 */)
{
  const double E = 1.0, nu = 0.3, lam = E * nu / ((1 + nu) * (1 - 2 * nu)), mu = E / (2 * (1 + nu));
  const double gp[2] = {0.5 - 0.5 / sqrt(3.0), 0.5 + 0.5 / sqrt(3.0)};
  const int ld = 8 * bs;
  for (int k = 0; k < ld * ld; k++) ke[k] = 0.0;
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int c = 0; c < 2; c++) {
    const double xi[3] = {gp[a], gp[b], gp[c]};
    double G[8][3];
    for (int n = 0; n < 8; n++) {
      const int o[3] = {n & 1, (n >> 1) & 1, (n >> 2) & 1};
      double f[3], df[3];
      for (int d = 0; d < 3; d++) { f[d] = o[d] ? xi[d] : 1.0 - xi[d]; df[d] = o[d] ? 1.0 : -1.0; }
      G[n][0] = df[0] * f[1] * f[2]; G[n][1] = f[0] * df[1] * f[2]; G[n][2] = f[0] * f[1] * df[2];
    }
    const double w = 1.0 / 8.0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
      double dot = G[i][0] * G[j][0] + G[i][1] * G[j][1] + G[i][2] * G[j][2];
      if (bs == 1) ke[i * ld + j] += w * dot;
      else
        for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++)
          ke[(i * 3 + p) * ld + j * 3 + q] += w * (lam * G[i][p] * G[j][q] + mu * G[i][q] * G[j][p] + (p == q ? mu * dot : 0.0));
    }
  }
}

struct SynthInfo { int kind, dim, cells[3], top; };
static std::map<uggpu_ctx *, SynthInfo> g_synth;

static SynthParams make_params(const SynthInfo &si, int level)
{
  SynthParams p;
  p.kind = si.kind;
  p.bs = si.kind == UGGPU_SYNTH_Q1_ELASTICITY ? 3 : 1;
  p.dim = si.dim;
  double h = 1.0 / (double)(si.cells[0] << level);
  p.hpow = si.dim == 3 ? h : 1.0;
  p.hvol = si.dim == 3 ? h * h * h : h * h;
  p.hhalf = 0.5 * h;
  return p;
}

extern "C" int uggpu_synth_hierarchy_part(uggpu_ctx *ctx, int kind, int nx, int ny, int nz, int top, int A,
                                          int px, int py, int pz, int rank, int64_t replicate_below)
{
  if (!ctx) return uggpu_fail(UGGPU_ERROR, "null context");
  if (kind != UGGPU_SYNTH_P1_SIMPLEX && kind != UGGPU_SYNTH_Q1_POISSON && kind != UGGPU_SYNTH_Q1_ELASTICITY && kind != UGGPU_SYNTH_P1_VARCOEF)
    return uggpu_fail(UGGPU_ERROR, "synthetic kind %d not implemented", kind);
  const bool simplex = kind == UGGPU_SYNTH_P1_SIMPLEX || kind == UGGPU_SYNTH_P1_VARCOEF;
  if (!simplex && nz <= 0) return uggpu_fail(UGGPU_ERROR, "Q1 hierarchies are generated in 3D only");
  if (nx < 1 || ny < 1 || nz < 0 || top < 0 || top >= UGGPU_MAX_LEVELS) return uggpu_fail(UGGPU_ERROR, "bad synthetic grid %dx%dx%d top %d", nx, ny, nz, top);
  const int dim = nz > 0 ? 3 : 2;
  if (px < 1 || py < 1 || pz < 1 || (dim == 2 && pz != 1)) return uggpu_fail(UGGPU_ERROR, "bad rank array %dx%dx%d", px, py, pz);
  const int nranks = px * py * pz;
  if (rank < 0 || rank >= nranks) return uggpu_fail(UGGPU_ERROR, "rank %d outside the %d-rank array", rank, nranks);
  if (nranks > 1 && (!ctx->comm || uggpu_comm_size(ctx) != nranks))
    return uggpu_fail(UGGPU_ERROR, "a %d-rank hierarchy needs uggpu_comm_init with %d ranks first", nranks, nranks);
  SynthInfo si;
  si.kind = kind; si.dim = dim; si.cells[0] = nx; si.cells[1] = ny; si.cells[2] = nz; si.top = top;
  const int P[3] = {px, py, pz};
  const int bs = kind == UGGPU_SYNTH_Q1_ELASTICITY ? 3 : 1;
  if (!simplex) {
    std::vector<double> ke((size_t)(8 * bs) * (8 * bs));
    hex_element_matrix(bs, ke.data());
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_ke, ke.data(), ke.size() * sizeof(double), 0, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  for (int l = 0; l <= top; l++) {
    int cells[3] = {nx << l, ny << l, dim == 3 ? nz << l : 0};
    int64_t n_global = (int64_t)(cells[0] + 1) * (cells[1] + 1) * (dim == 3 ? cells[2] + 1 : 1);
    const bool replicated = nranks == 1 || n_global <= replicate_below || l == 0;
    PartGrid *pg = new PartGrid();
    if (part_make(pg, dim, cells, P, rank, replicated ? 1 : 0)) { delete pg; return uggpu_fail(UGGPU_ERROR, "level %d: %dx%dx%d cells do not divide over %dx%dx%d ranks", l, cells[0], cells[1], cells[2], px, py, pz); }
    int64_t n64 = pg->n_own;
    if (n_global > 2147483000LL && replicated) { delete pg; return uggpu_fail(UGGPU_ERROR, "level %d would have %lld rows (int32 row indices)", l, (long long)n_global); }
    int n = (int)n64;
    UG_TRY(uggpu_level_create(ctx, l, n, bs));
    Level &L = ctx->lev[l];
    L.part = pg;
    L.partitioned = !replicated;
    L.nghost = replicated ? 0 : pg->n_ghost;
    L.n_global = n_global;
    UG_TRY(dalloc(ctx, &L.d_part, 1));
    CUDA_TRY(cudaMemcpyAsync(L.d_part, pg, sizeof(PartGrid), cudaMemcpyHostToDevice, ctx->stream));
    if (L.partitioned) {
      L.nnb = pg->nnb;
      L.nb_rank.assign(pg->nb_rank, pg->nb_rank + pg->nnb);
      L.nb_send_off.assign(pg->nb_send_off, pg->nb_send_off + pg->nnb + 1);
      L.nb_recv_off.assign(pg->nb_recv_off, pg->nb_recv_off + pg->nnb + 1);
      L.send_total = pg->nb_send_off[pg->nnb];
      UG_TRY(dalloc(ctx, &L.d_send_idx, (size_t)L.send_total));
      if (L.send_total > 0) {
        k_synth_sendidx<<<(L.send_total + 255) / 256, 256, 0, ctx->stream>>>(L.d_part, L.send_total, L.d_send_idx);
        KCHECK(ctx);
      }
    }
    SynthParams sp = make_params(si, l);
    if (n > 0) {
      k_synth_flags<<<(n + 255) / 256, 256, 0, ctx->stream>>>(L.d_part, n, bs, l == top, L.vclass, L.vnclass, L.ctl, L.skip);
      KCHECK(ctx);
    }
    SellMat m;
    UG_TRY(synth_sell<GEN_A>(ctx, sp, L.d_part, L.d_part, n, &m));
    UG_TRY(sell_update_diag(ctx, &m));
    UG_TRY(sell_share_values(ctx, &m));
    L.mats[A] = m;
    if (l > 0) {
      Level &Lc = ctx->lev[l - 1];
      UG_TRY(synth_sell<GEN_P>(ctx, sp, L.d_part, Lc.d_part, n, &L.P));
      UG_TRY(synth_sell<GEN_R>(ctx, sp, L.d_part, Lc.d_part, Lc.n, &L.R));
      UG_TRY(sell_compress_values(ctx, &L.P));
      UG_TRY(sell_compress_values(ctx, &L.R));
    }
  }
  ctx->fullrefinelevel = top;
  g_synth[ctx] = si;
  return 0;
}

extern "C" int uggpu_synth_hierarchy(uggpu_ctx *ctx, int kind, int nx, int ny, int nz, int top, int A)
{
  return uggpu_synth_hierarchy_part(ctx, kind, nx, ny, nz, top, A, 1, 1, 1, 0, 0);
}

extern "C" int uggpu_synth_rhs(uggpu_ctx *ctx, int level, int vec)
{
  auto it = g_synth.find(ctx);
  if (it == g_synth.end()) return uggpu_fail(UGGPU_ERROR, "context holds no synthetic hierarchy");
  Level *L = get_level(ctx, level);
  if (!L) return UGGPU_ERROR;
  UG_TRY(uggpu_vec_alloc(ctx, level, vec));
  if (L->n == 0) return 0;
  SynthParams sp = make_params(it->second, level);
  k_synth_rhs<<<(L->n + 255) / 256, 256, 0, ctx->stream>>>(sp, L->d_part, L->n, L->vecs[vec]);
  KCHECK(ctx);
  return 0;
}

extern "C" int64_t uggpu_level_n_global(uggpu_ctx *ctx, int level)
{
  Level *L = get_level(ctx, level);
  if (!L) return -1;
  return L->n_global > 0 ? L->n_global : (int64_t)L->n;
}

extern "C" int uggpu_level_is_partitioned(uggpu_ctx *ctx, int level)
{
  Level *L = get_level(ctx, level);
  return L ? (L->partitioned ? 1 : 0) : -1;
}

// ids[n] = global lexicographic index of every row this rank holds on `level`
extern "C" int uggpu_synth_global_ids(uggpu_ctx *ctx, int level, int64_t *ids)
{
  Level *L = get_level(ctx, level);
  if (!L || !L->d_part) return uggpu_fail(UGGPU_ERROR, "level %d is not synthetic", level);
  if (L->n == 0) return 0;
  int64_t *d = nullptr;
  UG_TRY(dalloc(ctx, &d, (size_t)L->n));
  k_synth_gids<<<(L->n + 255) / 256, 256, 0, ctx->stream>>>(L->d_part, L->n, d);
  KCHECK(ctx);
  CUDA_TRY(cudaMemcpyAsync(ids, d, (size_t)L->n * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  UG_TRY(dfree(ctx, d, (size_t)L->n));
  return 0;
}
