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

#include <cmath>
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

// ---- clusterAMG: MarkVanek amgtools.cc:254, CoarsenVanek :1960 with GenerateClusters :1864, IpPiecewiseConstant :3019 / IpVanek :3041 ------
// Aggregation: clusters are grown around seed vectors taken from bucket lists ordered by the number of strong connections from vectors
// that are still free (first pass: clusters of at least 2/3 of the average neighbourhood; second: leftovers join the smallest
// neighbouring cluster; third: the rest seeds clusters of any size).  One coarse vector per cluster, in the order of creation.
// cluster[v] = coarse vector of v, -1 for the vectors without strong connections (Dirichlet rows): they do not interpolate;
// seed[c] (or NULL) = the vector cluster c was started from.
// smooth = 0: piecewise constant interpolation; 1: Vanek's smoothed aggregation (one damped Jacobi step of the filtered matrix applied to
// the piecewise constant one; the cluster's own entry stays first in the row, the others follow in reverse order of creation).
extern "C" int uggpu_amg_vanek_host(int n, const int32_t *rowptr, const int32_t *col, const double *val, const uint32_t *skip, double theta, int smooth,
                                    int32_t *cluster, int32_t *seed, int32_t *p_rowptr, int32_t *p_col, double *p_w, int *n_coarse)
{
  if (n < 0 || !rowptr || !col || !val || !skip || !cluster || !p_rowptr || !p_col || !p_w || !n_coarse) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: null argument");
  const int64_t nnz = rowptr[n];
  for (int v = 0; v < n; v++)
    if (rowptr[v + 1] <= rowptr[v] || col[rowptr[v]] != v) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: row %d does not start with its diagonal entry", v);
  // ---- MarkVanek amgtools.cc:254-310, scalar (vcomp = 0)
  std::vector<uint8_t> strong((size_t)nnz, 0);
  for (int v = 0; v < n; v++) {
    if (skip[v]) continue;
    const double nii = fabs(val[rowptr[v]]);
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      const int w = col[e];
      if (skip[w] != 0) continue;
      const double njj = fabs(val[rowptr[w]]), nij = fabs(val[e]);
      if (nij >= theta * sqrt(nii * njj)) strong[e] = 1;
    }
  }
  std::vector<int32_t> adj((size_t)nnz, -1);
  for (int v = 0; v < n; v++)
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      const int w = col[e];
      for (int f = rowptr[w] + 1; f < rowptr[w + 1]; f++) if (col[f] == v) { adj[e] = f; break; }
    }
  // ---- CoarsenVanek :1960-2116
  std::vector<uint8_t> ccoarse(n, 0);
  std::vector<int> sin(n, 0), sout(n, 0);
  Lists Ls(n);
  int maxNeighbors = 0;
  long sumStrong = 0;
  for (int v = 0; v < n; v++) {                                   // CountStrongNeighbors :394
    int nb = 0, ns = 0;
    for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++) {
      if (strong[e]) { sumStrong++; sout[col[e]]++; ns++; }
      nb++;
    }
    if (nb > maxNeighbors) maxNeighbors = nb;
    sin[v] = ns;
  }
  const double avNosN = n > 0 ? (double)sumStrong / (double)n : 0.0;
  if (maxNeighbors > AMG_MAXNEIGHBORS) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: a row has %d neighbours, the coarsening handles %d (MAXNEIGHBORS)", maxNeighbors, AMG_MAXNEIGHBORS);
  const int nU = 2 * AMG_MAXNEIGHBORS + 1;
  std::vector<int> Ua(nU, -1), Ue(nU, -1);
  int Da = -1, De = -1;
  for (int v = 0; v < n; v++) {                                   // DistributeInitialList :354
    cluster[v] = -1;
    if (sin[v] == 0) Ls.add_end(Da, De, v);
    else { if (sout[v] >= nU) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: bucket overflow"); Ls.add_end(Ua[sout[v]], Ue[sout[v]], v); }
  }
  std::vector<int> csize;                                         // VINDEX(newVect): size of the cluster
  std::vector<int> cseed;                                         // the seed vector (its class becomes the next class of the coarse vector)
  int err = 0;
  auto lower_free_neighbours = [&](int m) {                       // "change the order for the neighbors": one strong connection from a free vector less
    for (int e = rowptr[m] + 1; e < rowptr[m + 1]; e++)
      if (strong[e]) {
        const int v2 = col[e];
        if (ccoarse[v2]) continue;
        int k = sout[v2];
        if (sin[v2] == 0 || k <= 0) { err = 1; return; }          // the reference would follow a NULL pointer here (a vector outside the bucket lists)
        Ls.eliminate(Ua[k], Ue[k], v2);
        sout[v2] = --k;
        Ls.add_end(Ua[k], Ue[k], v2);
      }
  };
  auto generate_clusters = [&](int minSize) {                     // GenerateClusters :1864-1958
    if (minSize < 0) minSize = 0;
    int i = AMG_MAXNEIGHBORS;
    std::vector<int> members;
    while (i >= minSize) {
      int a;
      while ((a = Ua[i]) != -1) {
        members.clear();
        Ls.eliminate(Ua[i], Ue[i], a);
        members.push_back(a);
        ccoarse[a] = 1;
        for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++) {
          const int e2 = adj[e];
          if (e2 < 0) { err = 2; return; }
          if (strong[e2]) {
            const int v2 = col[e];
            if (ccoarse[v2]) continue;                            // already belongs to a cluster
            const int k = sout[v2];
            Ls.eliminate(Ua[k], Ue[k], v2);
            members.push_back(v2);
            ccoarse[v2] = 1;
          }
        }
        const int c = (int)csize.size();
        csize.push_back((int)members.size());
        cseed.push_back(a);
        for (int m : members) { cluster[m] = c; lower_free_neighbours(m); if (err) return; }
      }
      i--;
    }
  };
  generate_clusters((int)((avNosN + 1.0) * 0.66 - 1.0));
  if (err) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: %s", err == 2 ? "G(A) is not symmetric" : "a strong connection leads to a vector without strong connections of its own");
  for (int i = 0; i < AMG_MAXNEIGHBORS; i++) {                    // second step :2041-2096
    int a = Ua[i];
    while (a != -1) {
      int minSize = 999, best = -1;
      for (int e = rowptr[a] + 1; e < rowptr[a + 1]; e++)
        if (strong[e]) {
          const int v2 = col[e];
          if (ccoarse[v2] && csize[cluster[v2]] < minSize) { minSize = csize[cluster[v2]]; best = cluster[v2]; }
        }
      if (best != -1) {
        ccoarse[a] = 1;
        lower_free_neighbours(a);
        if (err) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: a strong connection leads to a vector without strong connections of its own");
        Ls.eliminate(Ua[i], Ue[i], a);                            // its successor pointer stays valid
        cluster[a] = best;
        csize[best]++;
      }
      a = Ls.succ[a];
    }
  }
  generate_clusters(0);                                           // third pass
  if (err) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: %s", err == 2 ? "G(A) is not symmetric" : "a strong connection leads to a vector without strong connections of its own");
  const int nc = (int)csize.size();
  *n_coarse = nc;
  if (seed) for (int c = 0; c < nc; c++) seed[c] = cseed[c];
  // ---- interpolation
  int64_t z = 0;
  p_rowptr[0] = 0;
  std::vector<int32_t> oc; std::vector<double> ow;
  double factor = 0.0;
  for (int v = 0; v < n; v++) {
    if (cluster[v] < 0) {
      if (smooth && skip[v] == 0) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: vector %d without a cluster is not a Dirichlet vector (IpVanek needs its interpolation matrix)", v);
      p_rowptr[v + 1] = (int32_t)z; continue;
    }
    double own = 1.0;                                             // piecewise constant on the clusters :3019 / :3062
    oc.clear(); ow.clear();
    if (smooth && skip[v] == 0) {                                 // IpVanek :3066-3103: P := (I - 2/3 D_f^-1 A_strong) P, D_f the diagonal of the filtered matrix
      double sum = val[rowptr[v]];
      for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++)
        if (!strong[e] && skip[col[e]] == 0) sum += val[e];
      if (sum != 0.0) factor = 1.0 / sum;                         // BLOCK_INVERT: untouched when singular
      factor *= -0.666666666;
      for (int e = rowptr[v] + 1; e < rowptr[v + 1]; e++)
        if (strong[e]) {
          const int c2 = cluster[col[e]];
          if (c2 < 0) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_vanek_host: strong neighbour without a cluster");
          if (c2 == cluster[v]) { own += factor * val[e]; continue; }
          size_t k = 0;
          while (k < oc.size() && oc[k] != c2) k++;
          if (k == oc.size()) { oc.push_back(c2); ow.push_back(0.0); }
          ow[k] += factor * val[e];
        }
    }
    p_col[z] = cluster[v]; p_w[z] = own; z++;
    for (size_t k = oc.size(); k-- > 0;) { p_col[z] = oc[k]; p_w[z] = ow[k]; z++; }      // inserted right behind the cluster's entry: newest first
    p_rowptr[v + 1] = (int32_t)z;
  }
  return 0;
}

// the new level on the device from the interpolation rows: vectors' flags, by-matrix transfer stencils, Galerkin matrix
static int amg_build_level(uggpu_ctx *ctx, int level, int A, int n, int nc, const std::vector<uint8_t> &vclass, const std::vector<uint8_t> &cnclass,
                           const std::vector<uint32_t> &cskip, const std::vector<int32_t> &prp, const std::vector<int32_t> &pcol, const std::vector<double> &pw)
{
  std::vector<uint8_t> cclass((size_t)nc, 3), cctl((size_t)nc, 1);      // class 3, NEW_DEFECT set, FINE_GRID_DOF clear (amgtools.cc:585-600, :1905-1909)
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

extern "C" int uggpu_amg_coarsen_vanek(uggpu_ctx *ctx, int level, int A, double theta, int smooth, int *n_coarse)
{
  Level *L = get_level(ctx, level);
  SellMat *Af = get_mat(ctx, level, A);
  if (!L || !Af || !n_coarse) return UGGPU_DESC_MISMATCH;
  if (level < 1) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_coarsen_vanek: no room below level %d (levels are numbered from 0)", level);
  if (L->bs != 1) return uggpu_fail(UGGPU_BLOCK_TOO_LARGE, "uggpu_amg_coarsen_vanek: scalar equations only (block size %d)", L->bs);
  if (ctx->comm && L->partitioned) return uggpu_fail(UGGPU_ERROR, "uggpu_amg_coarsen_vanek runs on one GPU (level %d is partitioned)", level);
  const int n = L->n;
  const size_t nnz = (size_t)Af->nnz;
  std::vector<int32_t> rp((size_t)n + 1), col(nnz + 1), prp((size_t)n + 1), pcol(nnz + (size_t)n + 1), cluster((size_t)n + 1), seed((size_t)n + 1);
  std::vector<double> val(nnz + 1), pw(nnz + (size_t)n + 1);
  std::vector<uint8_t> vclass((size_t)n + 1);
  std::vector<uint32_t> skip((size_t)n + 1);
  UG_TRY(sell_to_host_csr(ctx, Af, rp.data(), col.data(), val.data()));
  UG_TRY(uggpu_level_get_flags(ctx, level, vclass.data(), nullptr, nullptr, skip.data()));
  int nc = 0;
  UG_TRY(uggpu_amg_vanek_host(n, rp.data(), col.data(), val.data(), skip.data(), theta, smooth, cluster.data(), seed.data(), prp.data(), pcol.data(), pw.data(), &nc));
  *n_coarse = nc;
  if (nc == 0 || nc == n) { *n_coarse = 0; return 0; }
  // GenerateClusters :1905-1909: next class = class of the seed vector; VECSKIP of a cluster's vector is never set
  std::vector<uint8_t> cnclass((size_t)nc);
  std::vector<uint32_t> cskip((size_t)nc, 0u);
  for (int c = 0; c < nc; c++) cnclass[c] = vclass[seed[c]];
  return amg_build_level(ctx, level, A, n, nc, vclass, cnclass, cskip, prp, pcol, pw);
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
  // the new level's vectors (GenerateNewGrid amgtools.cc:585-600): next class = class of the fine vector, VECSKIP inherited
  std::vector<uint8_t> cnclass((size_t)nc);
  std::vector<uint32_t> cskip((size_t)nc);
  for (int v = 0, k = 0; v < n; v++) if (coarse[v]) { cnclass[k] = vclass[v]; cskip[k] = skip[v]; k++; }
  return amg_build_level(ctx, level, A, n, nc, vclass, cnclass, cskip, prp, pcol, pw);
}
