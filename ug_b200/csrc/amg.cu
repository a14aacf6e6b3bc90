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

#include "amg_host.inc"

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
