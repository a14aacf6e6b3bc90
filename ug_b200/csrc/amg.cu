// amg.cu -- setup of ONE algebraic level below a given level (the AMG side): what one pass of the coarsening loop of
// AMGTransferPreProcess (np/procs/amgtransfer.cc:795-925) does for the reference's class `selectionAMG` configured as
//     $strongRel <theta> $C RugeStueben $I RugeStueben $CM Galerkin        (scalar equations)
//   MarkRelative            np/algebra/amgtools.cc:188   strong connections: -a_ij >= theta * max_k(-a_ik), Dirichlet rows / columns never
//   CoarsenRugeStueben      :684                         first pass (bucket lists by the number of strongly influenced undecided points),
//                                                        second pass (every strong F-F pair needs a common C point), GenerateNewGrid :538
//   IpRugeStueben           :2237                        interpolation weights of the F points from their strong C neighbours
//   AssembleGalerkinByMatrix transgrid.cc:1575           the coarse matrix, its pattern created by the product (galerkin.cu)
// The coarsening is a sequential graph algorithm whose result depends on the order of its list operations; it runs on the HOST on the flat
// matrix (uggpu_amg_rs_host: no device involved, pinned against the reference's levels by the CPU tests), as does the weight computation,
// which follows the reference's order of additions entry by entry.  The level it produces -- flags, by-matrix transfer stencils in the
// reference's list order, Galerkin matrix -- is created on the device; the cycle then runs on it like on a level the reference built.
// A setup path: one call per level, host time ~ the reference's own.
#include "uggpu_internal.h"

#include <vector>

#define AMG_MAXNEIGHBORS 128      // np/algebra/amgtools.h:52

namespace {

// doubly linked lists of vectors by index, with the reference's macros' semantics (amgtools.h:90-95): ELIMINATE trusts the caller about
// which list the item is in and only fixes that list's ends
struct Lists {
  std::vector<int> pred, succ;
  explicit Lists(int n) : pred(n, -1), succ(n, -1) {}
  void add_end(int &ls, int &le, int p) { pred[p] = le; succ[p] = -1; if (le != -1) succ[le] = p; else ls = p; le = p; }
  void eliminate(int &ls, int &le, int p)
  {
    if (pred[p] != -1) succ[pred[p]] = succ[p]; else ls = succ[p];
    if (succ[p] != -1) pred[succ[p]] = pred[p]; else le = pred[p];
  }
};

}  // namespace

extern "C" int uggpu_amg_rs_host(int n, const int32_t *rowptr, const int32_t *col, const double *val, const uint32_t *skip, double theta,
                                 uint8_t *coarse_out, int32_t *p_rowptr, int32_t *p_col, double *p_w, int *n_coarse)
{
  if (n < 0 || !rowptr || !col || !val || !skip || !coarse_out || !p_rowptr || !p_col || !p_w || !n_coarse) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: null argument");
  const int64_t nnz = rowptr[n];
  for (int v = 0; v < n; v++)
    if (rowptr[v + 1] <= rowptr[v] || col[rowptr[v]] != v) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: row %d does not start with its diagonal entry", v);
  // ---- MarkRelative amgtools.cc:188-252 (after UnmarkAll :108), scalar: the diagonal is never marked
  std::vector<uint8_t> strong((size_t)nnz, 0);
  for (int v = 0; v < n; v++) {
    if (skip[v]) continue;
    double s = 0.0;
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++)
      if (skip[col[e]] == 0) { const double nij = -val[e]; if (s < nij) s = nij; }
    const double threshold = s * theta;
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++)
      if (skip[col[e]] == 0 && -val[e] >= threshold) strong[e] = 1;
  }
  // MADJ: the entry (j, i) of every entry (i, j)
  std::vector<int32_t> adj((size_t)nnz, -1);
  for (int v = 0; v < n; v++)
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      const int w = col[e];
      for (int f = rowptr[w] + 1; f < rowptr[w + 1]; f++) if (col[f] == v) { adj[e] = f; break; }
    }
  // ---- CoarsenRugeStueben amgtools.cc:684-880
  std::vector<uint8_t> avcoarse(n, 0), avfine(n, 0), avtested(n, 0), used(n, 0);
  std::vector<int> sin(n, 0), sout(n, 0);
  Lists Ls(n);
  int maxNeighbors = 0;
  for (int v = 0; v < n; v++) {                                   // CountStrongNeighbors :394
    int nb = 0, ns = 0;
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      if (strong[e]) { sout[col[e]]++; ns++; }
      nb++;
    }
    if (nb > maxNeighbors) maxNeighbors = nb;
    sin[v] = ns;
  }
  if (maxNeighbors > AMG_MAXNEIGHBORS) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: a row has %d neighbours, the coarsening handles %d (MAXNEIGHBORS)", maxNeighbors, AMG_MAXNEIGHBORS);
  const int nU = 2 * maxNeighbors + 1;
  std::vector<int> Ua(nU, -1), Ue(nU, -1);
  int Ca = -1, Ce = -1, Fa = -1, Fe = -1, Ta = -1, Te = -1, Da = -1, De = -1;
  for (int v = 0; v < n; v++) {                                   // DistributeInitialList :354
    if (sin[v] == 0) { avfine[v] = 1; avtested[v] = 1; Ls.add_end(Da, De, v); }
    else { if (sout[v] >= nU) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: bucket overflow"); Ls.add_end(Ua[sout[v]], Ue[sout[v]], v); }
  }
  int i = maxNeighbors;
  while (i >= 0) {
    int a;
    while ((a = Ua[i]) != -1) {
      Ls.eliminate(Ua[i], Ue[i], a);
      Ls.add_end(Ca, Ce, a);
      avcoarse[a] = 1;
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++) {
        const int v2 = col[e];
        if (avfine[v2] || avcoarse[v2]) continue;
        const int e2 = adj[e];
        if (e2 < 0) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: G(A) is not symmetric");
        if (strong[e2]) {
          int k = sout[v2];
          Ls.eliminate(Ua[k], Ue[k], v2);
          Ls.add_end(Fa, Fe, v2);
          avfine[v2] = 1;
          for (int e3 = rowptr[v2] + 1; e3 < rowptr[v2 + 1]; e3++)
            if (strong[e3]) {
              const int v3 = col[e3];
              if (avfine[v3] || avcoarse[v3]) continue;
              k = sout[v3];
              Ls.eliminate(Ua[k], Ue[k], v3);
              k++;
              if (k >= nU) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: bucket overflow");
              if (k > i) i = k;
              sout[v3] = k;
              Ls.add_end(Ua[k], Ue[k], v3);
            }
        }
      }
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++)
        if (strong[e]) {
          const int v2 = col[e];
          if (avfine[v2] || avcoarse[v2]) continue;
          int k = sout[v2];
          Ls.eliminate(Ua[k], Ue[k], v2);
          --k;
          if (k < 0) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: bucket underflow");
          sout[v2] = k;
          Ls.add_end(Ua[k], Ue[k], v2);
        }
    }
    i--;
  }
  // second part: every F point's strong F neighbours must share a C point with it; otherwise one of the two becomes C
  {
    int a;
    while ((a = Fa) != -1) {
      Ls.eliminate(Fa, Fe, a);
      Ls.add_end(Ta, Te, a);
      avtested[a] = 1;
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++)
        if (strong[e] && avcoarse[col[e]]) used[col[e]] = 1;
      int testCoarse = -1;
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++)
        if (strong[e]) {
          const int v2 = col[e];
          if (used[v2]) continue;
          int flag = 0;
          for (int e2 = rowptr[v2] + 1; e2 < rowptr[v2 + 1]; e2++)
            if (strong[e2] && used[col[e2]]) { flag = 1; break; }
          if (flag == 0) {
            if (testCoarse == -1) { testCoarse = v2; used[v2] = 1; }
            else { testCoarse = a; break; }
          }
        }
      if (testCoarse != -1) {
        if (avtested[testCoarse]) Ls.eliminate(Ta, Te, testCoarse); else Ls.eliminate(Fa, Fe, testCoarse);
        Ls.add_end(Ca, Ce, testCoarse);
        avtested[testCoarse] = 0;
        avfine[testCoarse] = 0;
        if (skip[testCoarse]) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_rs_host: a Dirichlet vector would become a coarse point");      // assert :858
        avcoarse[testCoarse] = 1;
      }
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++) used[col[e]] = 0;
    }
  }
  // GenerateNewGrid :538: coarse vectors in the order of the fine list; nothing to do when all or none are coarse
  std::vector<int32_t> cindex(n, -1);
  int nc = 0;
  for (int v = 0; v < n; v++) { coarse_out[v] = avcoarse[v]; if (avcoarse[v]) cindex[v] = nc++; }
  *n_coarse = nc;
  if (nc == 0 || nc == n) { p_rowptr[0] = 0; for (int v = 0; v < n; v++) p_rowptr[v + 1] = 0; return 0; }
  // ---- IpRugeStueben :2237-2372, scalar.  tmp[k] is the reference's intermediate storage in the interpolation matrix of coarse point k;
  // the interpolation matrices of an F point are created in the order of its matrix list and CreateIMatrix inserts at the head
  // (gm/algebra.cc:7637), so the row lists them in reverse.
  std::vector<double> tmp(n, 0.0);
  std::vector<int32_t> rowc; std::vector<double> roww;
  double sumInv = 0.0, modDiagInv = 0.0;
  int64_t z = 0;
  p_rowptr[0] = 0;
  for (int v = 0; v < n; v++) {
    if (avcoarse[v]) { p_col[z] = cindex[v]; p_w[z] = 1.0; z++; p_rowptr[v + 1] = (int32_t)z; continue; }      // identity on the direct fathers :2365
    if (skip[v] == 0) {
      double modDiag = val[rowptr[v]];
      for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
        const int v2 = col[e];
        if (avcoarse[v2] && strong[e]) { used[v2] = 1; tmp[v2] = val[e]; }
        else if (!strong[e] && skip[v2] == 0) modDiag += val[e];                       // weak connections are lumped to the diagonal
      }
      for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++)
        if (strong[e]) {
          const int v2 = col[e];
          if (avcoarse[v2]) continue;
          double sum = 0.0;
          for (int e2 = rowptr[v2] + 1; e2 < rowptr[v2 + 1]; e2++) if (used[col[e2]]) sum += val[e2];
          if (sum != 0.0) sumInv = 1.0 / sum;                                          // BLOCK_INVERT amgtools.h:246: untouched when singular
          const double factor = val[e] * sumInv;
          for (int e2 = rowptr[v2] + 1; e2 < rowptr[v2 + 1]; e2++) if (used[col[e2]]) tmp[col[e2]] += factor * val[e2];
        }
      if (modDiag != 0.0) modDiagInv = 1.0 / modDiag;
      modDiagInv *= -1.0;
    }
    rowc.clear(); roww.clear();
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      const int v2 = col[e];
      if (used[v2]) { used[v2] = 0; rowc.push_back(cindex[v2]); roww.push_back(modDiagInv * tmp[v2]); }
    }
    for (size_t k = rowc.size(); k-- > 0;) { p_col[z] = rowc[k]; p_w[z] = roww[k]; z++; }
    p_rowptr[v + 1] = (int32_t)z;
  }
  return 0;
}

// One level: level-1 := coarsening of `level` with matrix A.  *n_coarse = 0 (and no level created) when the coarsening selects all or no
// vectors (GenerateNewGrid's "nothing to do", the reference's "error in coarsening").
extern "C" int uggpu_amg_coarsen_rs(uggpu_ctx *ctx, int level, int A, double theta, int *n_coarse)
{
  Level *L = get_level(ctx, level);
  SellMat *Af = get_mat(ctx, level, A);
  if (!L || !Af || !n_coarse) return UGGPU_DESC_MISMATCH;
  if (level < 1) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_coarsen_rs: no room below level %d (levels are numbered from 0)", level);
  if (L->bs != 1) return uggpu_fail(UGGPU_BLOCK_TOO_LARGE, "uggpu_amg_coarsen_rs: scalar equations only (block size %d)", L->bs);
  if (ctx->comm && L->partitioned) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_coarsen_rs runs on one GPU (level %d is partitioned)", level);
  const int n = L->n;
  const size_t nnz = (size_t)Af->nnz;
  std::vector<int32_t> rp((size_t)n + 1), col(nnz + 1), prp((size_t)n + 1), pcol(nnz + (size_t)n + 1);
  std::vector<double> val(nnz + 1), pw(nnz + (size_t)n + 1);
  std::vector<uint8_t> vclass((size_t)n + 1), coarse((size_t)n + 1);
  std::vector<uint32_t> skip((size_t)n + 1);
  UG_TRY(sell_to_host_csr(ctx, Af, rp.data(), col.data(), val.data()));
  UG_TRY(uggpu_level_get_flags(ctx, level, vclass.data(), nullptr, nullptr, skip.data()));
  int nc = 0;
  UG_TRY(uggpu_amg_rs_host(n, rp.data(), col.data(), val.data(), skip.data(), theta, coarse.data(), prp.data(), pcol.data(), pw.data(), &nc));
  *n_coarse = nc;
  if (nc == 0 || nc == n) { *n_coarse = 0; return 0; }
  // the new level's vectors (GenerateNewGrid amgtools.cc:585-600): class 3, next class = class of the fine vector, NEW_DEFECT set,
  // FINE_GRID_DOF clear, VECSKIP inherited
  std::vector<uint8_t> cclass((size_t)nc, 3), cnclass((size_t)nc), cctl((size_t)nc, 1);
  std::vector<uint32_t> cskip((size_t)nc);
  for (int v = 0, k = 0; v < n; v++) if (coarse[v]) { cnclass[k] = vclass[v]; cskip[k] = skip[v]; k++; }
  UG_TRY(uggpu_level_create(ctx, level - 1, nc, 1));
  UG_TRY(uggpu_level_set_flags(ctx, level - 1, cclass.data(), cnclass.data(), cctl.data(), cskip.data()));
  // R: the coarse rows list the contributions of the fine rows with VCLASS >= NEWDEF_CLASS in fine list order (RestrictByMatrix, transgrid.cc:1142)
  std::vector<int32_t> rrp((size_t)nc + 1, 0), rcol((size_t)prp[n] + 1);
  std::vector<double> rw((size_t)prp[n] + 1);
  for (int v = 0; v < n; v++) if (vclass[v] >= 2) for (int e = prp[v]; e < prp[v + 1]; e++) rrp[pcol[e] + 1]++;
  for (int k = 0; k < nc; k++) rrp[k + 1] += rrp[k];
  { std::vector<int32_t> fill(rrp.begin(), rrp.end() - 1);
    for (int v = 0; v < n; v++) if (vclass[v] >= 2) for (int e = prp[v]; e < prp[v + 1]; e++) { const int32_t pos = fill[pcol[e]]++; rcol[pos] = v; rw[pos] = pw[e]; } }
  UG_TRY(uggpu_transfer_set(ctx, level, prp.data(), pcol.data(), pw.data(), rrp.data(), rcol.data(), rw.data()));
  UG_TRY(uggpu_transfer_set_mode(ctx, level, UGGPU_TRANSFER_IMAT));
  return uggpu_galerkin(ctx, level, A);             // level-1 has no matrix A yet: pattern and values come from the product
}
