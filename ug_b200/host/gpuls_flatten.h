// gpuls_flatten.h -- PreProcess-time flattening of UG's linked VECTOR/MATRIX lists into the
// canonical flat layout of SURVEY.md 8(a') (BSR per level, per-row flags, P/R stencils).
//
// Host-side product code of the `gpuls` numproc family.  Compiled against the UG headers
// (twice: -D_2 / -D_3, like every UG source) by the application that links libug; it contains no
// CUDA and reaches the GPU only through include/uggpu.h.
//
// The traversal rule is the one the reference itself uses when it hands a grid to an array-based
// solver: np/amglib/amg_ug.cc:273-366 (index = position in FIRSTVECTOR->SUCCVC, row entries in
// VSTART->MNEXT order, diagonal first by gm/algebra.cc:1043-1048).
#ifndef GPULS_FLATTEN_H
#define GPULS_FLATTEN_H

#include <vector>
#include <cstdint>

#include "gm.h"
#include "np.h"
#include "udm.h"

namespace gpuls {

struct FlatLevel {
  int n = 0;                       // block rows = VECTORs in list order
  int bs = 0;                      // components per vector (NODEVEC)
  std::vector<int32_t> rowptr, col;
  std::vector<double>  val;        // nnz * bs*bs, row-major blocks
  std::vector<uint8_t> vclass, vnclass, ctl;
  std::vector<uint32_t> skip;
  // transfer to/from level-1 (empty on level 0)
  std::vector<int32_t> p_rowptr, p_col;  std::vector<double> p_w;
  std::vector<int32_t> r_rowptr, r_col;  std::vector<double> r_w;
  std::vector<int32_t> node_row;   // fine NODE list position -> row index (diagnostics / goldens)
};

// Number of components of `vd` in NODEVEC vectors; <=0 if the descriptor is not a pure nodal one.
int NodeComps(const NS_DIM_PREFIX VECDATA_DESC *vd);

// Writes VINDEX(v) = list position on `level` (same effect as l_setindex, np/algebra/ugiter.cc:169)
// and fills n, bs, flags.
int FlattenFlags(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX VECDATA_DESC *x, FlatLevel &out);
// Pattern + values of A on `level` (requires FlattenFlags first: uses VINDEX).
int FlattenMatrix(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX MATDATA_DESC *A, FlatLevel &out);
// Values only (pattern unchanged since FlattenMatrix).
int FlattenMatrixValues(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX MATDATA_DESC *A, std::vector<double> &val, int bs);
// Standard P and R between `level` and level-1 (requires FlattenFlags on both levels).
int FlattenTransfer(NS_DIM_PREFIX MULTIGRID *mg, int level, FlatLevel &out);

// IMAT mode (`transfer $M`, and what RestrictDefect/InterpolateCorrection use on levels < 1, np/procs/transfer.cc:724-760): P rows
// from the stored interpolation lists VISTART(v)->NEXT in list order, R rows in the order RestrictByMatrix_General
// (np/algebra/transgrid.cc:1113) scatters: fine VECTOR list order, entries of a vector in list order.  Only interpolation blocks
// of the form c*I (what CreateStandardNodeRestProl :2363 and the standard refinement store) are accepted: returns 7 otherwise,
// 8 when a vector has CRITBITs set.
int FlattenTransferIMAT(NS_DIM_PREFIX MULTIGRID *mg, int level, FlatLevel &out);

// Device-side assembly: the elements of `level` in FIRSTELEMENT->SUCCE order with the rows (VINDEX, FlattenFlags
// first) of their corner vectors in CORNER order, the vertex coordinates by row, and the VECSKIP words the element loop of
// np/procs/assemble.cc:657 would leave when every component of a boundary vertex met in a boundary element is Dirichlet
// (SetElementDirichletFlags, np/udm/disctools.cc:1763).
int FlattenElements(NS_DIM_PREFIX MULTIGRID *mg, int level, std::vector<int64_t> &elem_ptr, std::vector<int32_t> &elem_row, std::vector<double> &coord,
                    std::vector<uint32_t> &dirichlet_skip, int bs);
// inverse of FlattenMatrixValues / of the skip part of FlattenFlags: device results back into MVALUEs / VECSKIP
int ScatterMatrixValues(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX MATDATA_DESC *A, const std::vector<double> &val, int bs);
int ScatterSkip(NS_DIM_PREFIX MULTIGRID *mg, int level, const std::vector<uint32_t> &skip);

// VVALUE gather/scatter in list order: host[r*bs+i] <-> VVALUE(v, VD_CMP_OF_TYPE(vd,VTYPE(v),i))
void GatherVector(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX VECDATA_DESC *vd, int bs, double *host);
void ScatterVector(NS_DIM_PREFIX MULTIGRID *mg, int level, const NS_DIM_PREFIX VECDATA_DESC *vd, int bs, const double *host);

}  // namespace gpuls
#endif
