/* uggpu.h -- C-ABI of libuggpu.so, the B200 (sm_100a) device layer behind the
 * `gpuls` numproc family (ug_b200/host/gpuls_np.cc).
 *
 * No UG type crosses this boundary: plain pointers, ints and doubles only.
 * Every function returns 0 on success (NUM_OK, np/np.h:66) and a non-zero code on
 * failure; uggpu_last_error() then describes it.  Error codes reuse the reference's
 * NUM_* values (np/np.h:65-75) where one applies.  There is NO CPU fallback: if no
 * CUDA device is usable, uggpu_ctx_create fails.
 *
 * Data model (SURVEY.md 8a'): a context holds a hierarchy of levels 0..top.  A level
 * has n block rows (one per UG VECTOR in FIRSTVECTOR->SUCCVC order), block size bs
 * (components per vector, 1..UGGPU_MAX_BS), per-row flags copied from the VECTOR
 * control words, any number of matrices (one per MATDATA_DESC, "mat" handle) in
 * BSR with the entries of a row in VSTART->MNEXT order (diagonal first), any number
 * of vectors (one per VECDATA_DESC, "vec" handle, dense double[n*bs]), and the
 * standard prolongation P (rows = this level, cols = level-1) together with the
 * restriction R (rows = level-1, cols = this level, entries in fine NODE list order).
 *
 * Every operation is asynchronous on the context's stream except the ones that
 * return host values (reductions, downloads, solve), which synchronise before
 * returning -- the reference's callers read VVALUEs / LRESULT immediately.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference root).
 */
#ifndef UGGPU_H
#define UGGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UGGPU_MAX_BS      3    /* components per vector handled by the block kernels          */
#define UGGPU_MAX_LEVELS  64   /* 2 * MAXLEVEL (gm/gm.h): the geometric levels and, below them, the algebraic levels of an AMG transfer */
#define UGGPU_MAX_COMP    40   /* MAX_VEC_COMP, np/udm/udm.h: length of a VEC_SCALAR          */

/* loop modes, np/np.h:184-185 */
#define UGGPU_ON_SURFACE  (-1)
#define UGGPU_ALL_VECTORS 0

/* bits of the per-row `ctl` byte (gm/gm.h:2264-2279 FINE_GRID_DOF / NEW_DEFECT) */
#define UGGPU_CTL_NEW_DEFECT    1u
#define UGGPU_CTL_FINE_GRID_DOF 2u

/* return codes (np/np.h:65-75) */
#define UGGPU_OK               0
#define UGGPU_OUT_OF_MEM       1
#define UGGPU_DESC_MISMATCH    3
#define UGGPU_BLOCK_TOO_LARGE  4
#define UGGPU_SMALL_DIAG       6
#define UGGPU_NO_COARSER_GRID  7
#define UGGPU_ERROR            9
#define UGGPU_CUDA_ERROR      20

typedef struct uggpu_ctx uggpu_ctx;

/* ---- context ------------------------------------------------------------------------ */
int  uggpu_ctx_create(int device, uggpu_ctx **out);
int  uggpu_ctx_destroy(uggpu_ctx *ctx);
const char *uggpu_last_error(void);
int  uggpu_sync(uggpu_ctx *ctx);
/* the context's CUDA stream (cudaStream_t as void*), so that callers can time with events on it */
int  uggpu_stream(uggpu_ctx *ctx, void **stream);
/* FULLREFINELEVEL(mg) (gm/gm.h) used by the ON_SURFACE loops, np/algebra/vecloop.ct:22 */
int  uggpu_set_fullrefinelevel(uggpu_ctx *ctx, int level);
/* number of kernels launched by this context since creation (bench `gpu_launches`) */
int64_t uggpu_launch_count(uggpu_ctx *ctx);
/* bytes of device memory currently held by the context */
int64_t uggpu_device_bytes(uggpu_ctx *ctx);

/* ---- per-kernel timing (CUDA events on the context's stream around every launch while enabled) -------------
 * kinds; level/kind -1 = all.  alg_bytes = algorithmic (compulsory) bytes of the matching launches, SURVEY.md 8(d). */
#define UGGPU_K_SMOOTH      0   /* fused smoothing step (dmatmul_minus + dadd + next l_jac [+ x update, norm]) */
#define UGGPU_K_JAC         1
#define UGGPU_K_RESTRICT    2
#define UGGPU_K_INTERPOLATE 3
#define UGGPU_K_VECOP       4
#define UGGPU_K_REDUCE      5
#define UGGPU_K_DMATMUL     6
#define UGGPU_K_BASE        7
#define UGGPU_K_TRISOLVE    8   /* level-scheduled triangular solve of the Gauss-Seidel family */
#define UGGPU_K_HALO        9   /* stand-alone halo exchange of a partitioned level (pushes fused into a producing kernel are part of that kernel) */
#define UGGPU_K_ALLREDUCE  10   /* ncclAllReduce of norm / dot partial sums and of the gathered coarse defect */
#define UGGPU_K_ASSEMBLE   11   /* element-loop assembly of one level (uggpu_assemble) */
#define UGGPU_K_GALERKIN   12   /* Galerkin product of one level (uggpu_galerkin: the product kernel, without the sort of the transposed stencil) */
int uggpu_prof_enable(uggpu_ctx *ctx, int on);   /* also clears the records */
int uggpu_prof_summary(uggpu_ctx *ctx, int kind, int level, int64_t *launches, double *ms, double *alg_bytes);

/* ---- hierarchy upload = "PreProcess flattens VECTOR/MATRIX lists" ------------------------
 * (the reference's own precedent: np/amglib/amg_ug.cc:207-390 AMGSolverPreProcess)       */
int uggpu_level_create(uggpu_ctx *ctx, int level, int n, int bs);
int uggpu_level_destroy(uggpu_ctx *ctx, int level);
int uggpu_level_n(uggpu_ctx *ctx, int level);
int uggpu_level_bs(uggpu_ctx *ctx, int level);
/* vclass/vnclass: VCLASS/VNCLASS 0..3 (gm/gm.h:2224-2232); ctl: UGGPU_CTL_*; skip: VECSKIP bits.
 * NULL pointers select the defaults of a uniformly refined grid: class 3, nclass 3 (0 on the
 * top level is irrelevant to the kernels), ctl = both bits, skip = 0. */
int uggpu_level_set_flags(uggpu_ctx *ctx, int level, const uint8_t *vclass, const uint8_t *vnclass,
                          const uint8_t *ctl, const uint32_t *skip);
/* rowptr[n+1], col[nnz] (int32, 0-based), val[nnz*bs*bs] row-major blocks; host pointers. */
int uggpu_level_get_flags(uggpu_ctx *ctx, int level, uint8_t *vclass, uint8_t *vnclass, uint8_t *ctl, uint32_t *skip);
int uggpu_mat_set(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col,
                  const double *val);
int uggpu_mat_set_values(uggpu_ctx *ctx, int level, int mat, const double *val);
/* pattern only, all values 0 (what creatematrix + dmatset(A, 0) leave): the matrix a device-side assembly fills */
int uggpu_mat_set_pattern(uggpu_ctx *ctx, int level, int mat, const int32_t *rowptr, const int32_t *col);
int uggpu_mat_get(uggpu_ctx *ctx, int level, int mat, int32_t *rowptr, int32_t *col, double *val);
int64_t uggpu_mat_nnz(uggpu_ctx *ctx, int level, int mat);
/* entries actually stored on the device (SELL-32 slices padded to their longest row) */
int64_t uggpu_mat_padded_nnz(uggpu_ctx *ctx, int level, int mat);
/* int32 column words one pass over the matrix reads: slices whose rows all have the same column distances store one
 * word per slice column instead of one per entry (lossless; uggpu_mat_get returns the original indices) */
int64_t uggpu_mat_col_words(uggpu_ctx *ctx, int level, int mat);
/* Entries whose VALUES a pass over the matrix fetches from HBM: slices of rows with uniform column distances AND bit-identical
 * values per slice column (interior rows of a constant-coefficient operator on a structured / uniformly refined grid) read one
 * shared table instead of their own values (lossless, verified on the device; UGGPU_NO_SHARED_VALUES=1 switches it off).
 * Equals uggpu_mat_nnz when nothing is shared. */
int64_t uggpu_mat_val_entries(uggpu_ctx *ctx, int level, int mat);
/* Scalar matrices: slices (of 32 rows) that carry the matrix' DOMINANT stencil -- one pair of distance and value tables used by
 * more than half of the slices; 0 when there is none.  Such matrices run the stencil variant of the fused smoothing kernel
 * (tables in the kernel's constant bank; same arithmetic, same results; UGGPU_NO_STENCIL=1 switches it off). */
int64_t uggpu_mat_stencil_slices(uggpu_ctx *ctx, int level, int mat);
/* Bytes of MATRIX data one SpMV-type pass over the matrix fetches from HBM in its current storage form (values that are read, column
 * words, row lengths / slice offsets; for matrices served by the stencil-rows / exception-rows kernels: the row mask and the packed
 * exception rows).  What the roofline figures count for the matrix; SURVEY.md 8(d)'s model is (8 b^2 + 4) * nnz + 4 (n + 1). */
double uggpu_mat_pass_bytes(uggpu_ctx *ctx, int level, int mat);
int uggpu_mat_free(uggpu_ctx *ctx, int level, int mat);
/* Standard (geometric) transfer stencils between `level` and level-1 (np/algebra/transgrid.cc:117-336):
 * P: p_rowptr[n_fine+1], p_col (coarse row), p_w (GNs weight, zeros dropped, corner order);
 * R: r_rowptr[n_coarse+1], r_col (fine row), r_w, entries in FIRSTNODE(fine)->SUCCN order and only
 *    for fine vectors with VCLASS >= NEWDEF_CLASS (transgrid.cc:153). */
int uggpu_transfer_set(uggpu_ctx *ctx, int level,
                       const int32_t *p_rowptr, const int32_t *p_col, const double *p_w,
                       const int32_t *r_rowptr, const int32_t *r_col, const double *r_w);

/* Transfer mode of `level` (np/procs/transfer.cc:553-573): UGGPU_TRANSFER_STANDARD (default) = StandardRestrict /
 * StandardInterpolateCorrection (the damping enters every term, transgrid.cc:164,305); UGGPU_TRANSFER_IMAT = `transfer $M`:
 * RestrictByMatrix / InterpolateCorrectionByMatrix (transgrid.cc:1113,1292) on stencils taken from the stored interpolation
 * matrices -- P rows in VISTART->NEXT order, R rows in fine VECTOR list order -- where the sums are formed without the damping
 * and the finished vector is scaled afterwards if a factor differs from 1. */
#define UGGPU_TRANSFER_STANDARD 0
#define UGGPU_TRANSFER_IMAT     1
int uggpu_transfer_set_mode(uggpu_ctx *ctx, int level, int mode);

/* which: 0 = P, 1 = R; any output pointer may be NULL */
int uggpu_transfer_get(uggpu_ctx *ctx, int level, int which, int32_t *rowptr, int32_t *col, double *w);
int64_t uggpu_transfer_nnz(uggpu_ctx *ctx, int level, int which);

/* ---- vectors (VECDATA_DESC on one level) ------------------------------------------------ */
int uggpu_vec_alloc(uggpu_ctx *ctx, int level, int vec);          /* AllocVDFromVD, np/udm/udm.h:476 */
int uggpu_vec_free(uggpu_ctx *ctx, int level, int vec);           /* FreeVD, udm.h:520 */
int uggpu_vec_upload(uggpu_ctx *ctx, int level, int vec, const double *host);   /* n*bs doubles */
int uggpu_vec_download(uggpu_ctx *ctx, int level, int vec, double *host);
/* Upload on a second stream: returns at once, the copy runs behind the work already enqueued and the first later operation
 * that touches the vector waits for it; `host` (pinned for a truly asynchronous copy) must stay valid until then.  With the
 * fused cycle the iterate x of uggpu_ls_solve is touched by the last kernel of a cycle only, so its upload hides behind it. */
int uggpu_vec_upload_async(uggpu_ctx *ctx, int level, int vec, const double *host);
/* raw device pointer of a vector (for callers that already hold device data, e.g. the bench) */
int uggpu_vec_devptr(uggpu_ctx *ctx, int level, int vec, void **dptr);

/* ---- BLAS level 1, np/np.h:190-226, np/algebra/ugblas.cc:2291-3251 ----------------------------- */
int uggpu_dset     (uggpu_ctx*, int fl, int tl, int mode, int x, double a);
int uggpu_dcopy    (uggpu_ctx*, int fl, int tl, int mode, int x, int y);
int uggpu_dscal    (uggpu_ctx*, int fl, int tl, int mode, int x, double a);
int uggpu_dscalx   (uggpu_ctx*, int fl, int tl, int mode, int x, const double *a /* [bs] */);
int uggpu_dadd     (uggpu_ctx*, int fl, int tl, int mode, int x, int y);
int uggpu_dsub     (uggpu_ctx*, int fl, int tl, int mode, int x, int y);
int uggpu_dminusadd(uggpu_ctx*, int fl, int tl, int mode, int x, int y);
int uggpu_daxpy    (uggpu_ctx*, int fl, int tl, int mode, int x, double a, int y);
int uggpu_daxpyx   (uggpu_ctx*, int fl, int tl, int mode, int x, const double *a, int y);
int uggpu_ddot     (uggpu_ctx*, int fl, int tl, int mode, int x, int y, double *a);
int uggpu_ddotx    (uggpu_ctx*, int fl, int tl, int mode, int x, int y, double *a /* [bs] */);
int uggpu_dnrm2    (uggpu_ctx*, int fl, int tl, int mode, int x, double *a);
int uggpu_dnrm2x   (uggpu_ctx*, int fl, int tl, int mode, int x, double *a /* [bs] */);

/* ---- BLAS level 2, np/np.h:238-243, ugblas.cc:3782-4042 ------------------------------------------ */
int uggpu_dmatmul      (uggpu_ctx*, int fl, int tl, int mode, int x, int M, int y);  /* x  = M y */
int uggpu_dmatmul_add  (uggpu_ctx*, int fl, int tl, int mode, int x, int M, int y);  /* x += M y */
int uggpu_dmatmul_minus(uggpu_ctx*, int fl, int tl, int mode, int x, int M, int y);  /* x -= M y */

/* ---- smoother, np/np.h:430 l_jac (np/algebra/ugiter.cc:271-335) ------------------------------------ */
int uggpu_l_jac(uggpu_ctx*, int level, int v, int M, int d);
/* Smoother() of np/procs/iter.cc:817-842 with Step = JacobiStep: x = damp * Diag(A)^-1 b ; b -= A x */
int uggpu_jac_smooth(uggpu_ctx*, int level, int x, int b, int A, const double *damp /* [bs] */);

/* ---- Gauss-Seidel family, np/np.h:431-441 (np/algebra/ugiter.cc:412 l_lgs, :735 l_ugs, :1343 l_lsor, :1563 l_usor) --------
 * Triangular solves in VINDEX (= row) order: v = (D+L)^-1 d / (D+U)^-1 d, with relaxation omega[bs] for the sor variants
 * (scalar rows: omega*(d-sum)/diag; block rows: SolveSmallBlock, then v_i *= omega_i).  Rows with VCLASS < ACTIVE_CLASS get 0
 * and are skipped as columns.  Results are bit-identical to the reference: rows are scheduled by dependency level, every row
 * still adds the reference's terms in VSTART->MNEXT order.  uggpu_gs_preprocess builds the schedule of matrix M on `level`
 * (GSPreProcess iter.cc:1003: l_setindex); the solves build it on first use if it is missing.  The pattern must be
 * structurally symmetric (UG's CONNECTIONs are MATRIX pairs, gm/gm.h:653).  One GPU only. */
int uggpu_gs_preprocess(uggpu_ctx*, int level, int M);
int uggpu_gs_levels(uggpu_ctx*, int level, int M, int *lower, int *upper);      /* dependency levels of the two schedules (0 = not built) */
int uggpu_l_lgs (uggpu_ctx*, int level, int v, int M, int d);
int uggpu_l_ugs (uggpu_ctx*, int level, int v, int M, int d);
int uggpu_l_lsor(uggpu_ctx*, int level, int v, int M, int d, const double *omega /* [bs] */);
int uggpu_l_usor(uggpu_ctx*, int level, int v, int M, int d, const double *omega /* [bs] */);
/* smoother classes of np/procs/iter.cc (iter.cc:10343-10366) */
#define UGGPU_SM_JAC 0   /* jac: Smoother :817 + JacobiStep :911 */
#define UGGPU_SM_GS  1   /* gs:  Smoother :817 + GSStep :1039    */
#define UGGPU_SM_SGS 2   /* sgs: SGSSmoother :1392               */
#define UGGPU_SM_SOR 3   /* sor: SORSmoother :4786 + SORStep :4744 (damp acts as omega inside l_lsor) */
#define UGGPU_SM_ILU 4   /* ilu: Smoother :817 + ILUStep :5478 (l_luiter on the decomposition made by ILUPreProcess :5444) */

/* ---- ILU, np/np.h:236,456,466 ------------------------------------------------------------------------------
 * dmatcopy (np/algebra/ugblas.cc, np.h:236): M = A on levels fl..tl, mode UGGPU_ALL_VECTORS only.  A missing M is created
 * with the pattern of A (what AllocMDFromMD + dmatcopy do in ILUPreProcess iter.cc:5457-5461); an existing M must have it.
 * l_ilubthdecomp (np/algebra/ugiter.cc:2252) as class `ilu` calls it -- beta[bs] (NULL: no diagonal modification), no threshold,
 * no rest vector, hence no new connections: incomplete decomposition of M on its own pattern, in place; the diagonal blocks are
 * stored inverted (StoreInverse :139).  Rows are processed by dependency level of the lower triangle, every row receiving the
 * reference's updates in the reference's order (pivot rows ascending, their entries in VSTART->MNEXT order): bit-identical to
 * the sequential elimination.  Returns UGGPU_SMALL_DIAG when a diagonal (block) cannot be inverted.  It also (re)builds the
 * two triangular-solve schedules of M.  One GPU only.
 * l_luiter (ugiter.cc:4444): v = U^-1 L^-1 d with Diag(L) = I and the stored inverse diagonal of U; inactive rows get 0. */
int uggpu_dmatcopy(uggpu_ctx*, int fl, int tl, int mode, int M, int A);
int uggpu_l_ilubthdecomp(uggpu_ctx*, int level, int M, const double *beta /* [bs] or NULL */);
int uggpu_l_luiter(uggpu_ctx*, int level, int v, int M, int d);

/* One smoothing step of class `kind` in defect-correction form: on entry b = defect, on exit x = correction and b = new
 * defect.  tmp: sgs -- handle of a work VECTOR (NP_SGS_t iter.cc:1386); ilu -- handle of the decomposed MATRIX
 * (NP_SMOOTHER.L iter.cc:5459); ignored by the other classes. */
int uggpu_smooth(uggpu_ctx*, int level, int kind, int x, int b, int A, const double *damp /* [bs] */, int tmp);

/* ---- grid transfer, np/np.h:475-489 (np/algebra/transgrid.cc:462,529) ------------------------------- */
/* StandardRestrict(GRID_ON_LEVEL(level), to, from, damp): fine `level` -> level-1 */
int uggpu_restrict(uggpu_ctx*, int level, int to, int from, const double *damp /* [bs] */);
/* StandardInterpolateCorrection(GRID_ON_LEVEL(level), to, from, damp): level-1 -> fine `level` */
int uggpu_interpolate_correction(uggpu_ctx*, int level, int to, int from, const double *damp);

/* ---- Galerkin coarse-grid operator, np/np.h:540, np/algebra/transgrid.cc:1575 ------------------------------
 * AssembleGalerkinByMatrix(GRID_ON_LEVEL(level), A, 0) after dmatset(level-1, level-1, ALL_VECTORS, A, 0.0), i.e. what `npcheck $G`
 * does (np/algebra/npcheck.cc:375-379): matrix A of level-1 := P^T A_level P on the interpolation stencils of `level`, every
 * coarse entry receiving the reference's terms in the reference's order ((m*im)*jm over the fine rows in list order, their entries
 * in list order, the interpolation entries of the neighbour in list order) -- bit-identical values.  Connections the product needs
 * and level-1 lacks are created like the reference does (CreateExtraConnection, transgrid.cc:1615: second place of both rows' lists):
 * when A does not exist on level-1 (a fresh algebraic level, np/procs/amgtransfer.cc:915) its pattern comes from the product alone,
 * diagonal entries first; when the product leaves an existing pattern that pattern grows.  The symbolic step runs on the host
 * (uggpu_galerkin_pattern), the numeric one on the device.  One GPU only. */
int uggpu_galerkin(uggpu_ctx*, int level, int A);
/* Host only (no device): the pattern of the coarse level after the product -- per row the diagonal, the connections the product creates
 * in REVERSE order of creation (creation order = first term (iv, jv) or (jv, iv) in the reference's traversal; gm/algebra.cc:1051-1078),
 * then the off-diagonal entries of the start pattern.  start_* == NULL: one diagonal entry per row.  out_col == NULL: row pointers only. */
int uggpu_galerkin_pattern(int nf, int nc, const int32_t *a_rowptr, const int32_t *a_col, const int32_t *p_rowptr, const int32_t *p_col,
                           const int32_t *start_rowptr, const int32_t *start_col, int32_t *out_rowptr, int32_t *out_col);

/* ---- setup of algebraic levels (the AMG side), np/procs/amgtransfer.cc:795-925 with np/algebra/amgtools.cc ---------------
 * One pass of the reference's coarsening loop for `selectionAMG $strongRel <theta> $C RugeStueben $I RugeStueben $CM Galerkin` on scalar
 * equations: MarkRelative (amgtools.cc:188), CoarsenRugeStueben (:684) + GenerateNewGrid (:538), IpRugeStueben (:2237), then the Galerkin
 * matrix (uggpu_galerkin; its pattern is created by the product).  level-1 is created: flags of its vectors, by-matrix transfer stencils in
 * the reference's list order, matrix A -- identical to the level the reference builds, bit for bit, so a cycle over it gives the
 * reference's results.  The coarsening and the weights are computed on the host from the downloaded matrix (sequential list algorithm;
 * uggpu_amg_rs_host is that half alone, no device involved); *n_coarse = 0 and no level when all or no vectors would be coarse (the
 * reference's "error in coarsening").  Levels are numbered from 0: the caller places the finest level high enough.  One GPU only. */
int uggpu_amg_coarsen_rs(uggpu_ctx*, int level, int A, double theta, int *n_coarse);
/* coarse[n]: 1 for the vectors that become coarse points (they are numbered in list order); P in by-matrix order: p_rowptr[n+1],
 * p_col / p_w with room for nnz + n entries.  All rows of P are empty when *n_coarse is 0 or n. */
int uggpu_amg_rs_host(int n, const int32_t *rowptr, const int32_t *col, const double *val, const uint32_t *skip, double theta,
                      uint8_t *coarse, int32_t *p_rowptr, int32_t *p_col, double *p_w, int *n_coarse);

/* The same for `clusterAMG $strongVanek <theta> $C VanekNeuss $I {PiecewiseConstant | Vanek} $CM Galerkin` (scalar): MarkVanek (amgtools.cc:254),
 * CoarsenVanek + GenerateClusters (:1960, :1864: aggregation in three passes), IpPiecewiseConstant (:3019) for smooth = 0 or IpVanek (:3041,
 * smoothed aggregation) for smooth = 1.  One coarse vector per cluster in the order of creation.  uggpu_amg_vanek_host: cluster[n] (-1: no
 * cluster, the vector does not interpolate), seed[n] (first n_coarse entries: the vector each cluster was started from; may be NULL). */
int uggpu_amg_coarsen_vanek(uggpu_ctx*, int level, int A, double theta, int smooth, int *n_coarse);
int uggpu_amg_vanek_host(int n, const int32_t *rowptr, const int32_t *col, const double *val, const uint32_t *skip, double theta, int smooth,
                         int32_t *cluster, int32_t *seed, int32_t *p_rowptr, int32_t *p_col, double *p_w, int *n_coarse);

/* ---- element-loop assembly on the device, np/procs/assemble.h:225 NP_LOCAL_ASSEMBLE, np/procs/assemble.cc:657 ------
 * One level of LocalAssemble (assemble.cc:671-697) followed by that level's share of NPLocalAssemblePostMatrix (:624): b = 0, A = 0,
 * VECSKIP cleared; for the elements in list order the local defect and the local matrix (summed over the quadrature points) are added
 * to the vectors' / connections' values (GetElementVVMPtrs np/udm/disctools.cc:1113: corner order, components row-major per block
 * pair); VECSKIP := skip (SetElementDirichletFlags :1763 -- which components are Dirichlet is the application's decision); then
 * AssembleDirichletBoundary (disctools.cc:1837): for every component with its skip bit set  b = x, the row of the diagonal block becomes
 * the unit row and the row of every other block of the vector 0.  x is only read (the caller has set its Dirichlet values, the job of
 * AssembleLocal in the reference).  The element kernel -- application code in UG -- is built in:
 *   UGGPU_FE_POISSON     bs = 1:   int coef_e grad(phi_i).grad(phi_j),  rhs int source[0] phi_i
 *   UGGPU_FE_ELASTICITY  bs = dim: isotropic linear elasticity (E, nu; lambda = E nu / ((1+nu)(1-2nu)), mu = E / (2(1+nu))), scaled by coef_e
 * on simplices (dim+1 corners, centroid rule) and tensor elements (2^dim corners in UG's corner numbering, 2-point Gauss per direction).
 * elem_ptr[nelem+1] / elem_row: rows of the elements' corner vectors in CORNER order, elements in FIRSTELEMENT->SUCCE order; coef[nelem]
 * (NULL: 1); coord[n*dim] by row; skip[n] (NULL: none); host pointers.  Matrix A must exist with its pattern (uggpu_mat_set /
 * uggpu_mat_set_pattern), vectors x and b must exist.  Every value receives the reference's terms in the reference's order (one thread
 * per row gathers its elements in list order): bit-identical to the sequential scatter loop.  One GPU only. */
#define UGGPU_FE_POISSON    0
#define UGGPU_FE_ELASTICITY 1
typedef struct uggpu_fe_cfg {
  int    problem;                    /* UGGPU_FE_*                                   */
  int    dim;                        /* 2 | 3                                        */
  double E, nu;                      /* elasticity                                   */
  double source[UGGPU_MAX_BS];       /* right-hand side density per component        */
} uggpu_fe_cfg;
int uggpu_assemble(uggpu_ctx*, int level, int x, int b, int A, const uggpu_fe_cfg *cfg, int64_t nelem, const int64_t *elem_ptr,
                   const int32_t *elem_row, const double *coef, const double *coord, const uint32_t *skip);

/* ---- savedata / loaddata for device vectors, np/udm/data_io.cc:650 SaveData / :408 LoadData -----------------------------
 * The reference's data files: a header (np/udm/dio.cc:338 Write_DT_General; DIO_GENERAL np/udm/dio.h) and one record per NODE, in the
 * order of the node IDs over all levels, holding the components of the saved VECDATA_DESCs side by side; modes "asc" and "bin"
 * (low/bio.cc; "xdr" is not offered).  Files written here are byte-identical to the reference's and either side reads the other's.
 * The caller supplies what only the grid manager knows: for node ID i the level id_level[i] and the row id_row[i] of its vector, and the
 * header fields.  uggpu_savedata gathers the vectors `vec[0..nvd)` of all levels on the device into file order, copies the body down
 * once and writes it; uggpu_loaddata parses a file, uploads the body once and scatters it into vec[0..nvd) (vec[i] < 0, or fewer
 * descriptors than the file holds: those values are skipped, data_io.cc:515-523; component counts must match, :521).
 * uggpu_data_write / uggpu_data_read are the host-only halves (format only, no device): data[nnode * sum(ncomp)] in file order. */
typedef struct uggpu_data_general {      /* DIO_GENERAL, np/udm/dio.h                                                        */
  const char *ident;                     /* string variable :IDENTIFICATION ("---" when unset, data_io.cc:750)               */
  const char *mgfile;                    /* multigrid file the data belongs to ("saved_without_mg", data_io.cc:747)          */
  double time, dt, ndt;                  /* -1 when the file carries no time step number (data_io.cc:758-762)                */
  int    nparfiles, me, magic_cookie;    /* procs, me, MG_MAGIC_COOKIE(theMG)                                                */
} uggpu_data_general;
int uggpu_data_write(const char *filename, const char *type /* "asc" | "bin" */, const uggpu_data_general *g, int nvd, const int *ncomp,
                     const char *const *vdname, const char *const *compnames, int64_t nnode, const double *data);
/* data == NULL: header only.  g->ident / g->mgfile point into storage that stays valid until the next call on the same thread. */
int uggpu_data_read(const char *filename, uggpu_data_general *g, int *nvd, int *ncomp /* [ncomp_cap] or NULL */, int ncomp_cap,
                    int64_t *ndata, double *data, int64_t data_cap);
int uggpu_savedata(uggpu_ctx *ctx, const char *filename, const char *type, const uggpu_data_general *g, int nvd, const int *vec,
                   const char *const *vdname, const char *const *compnames, int64_t nnode, const int32_t *id_level, const int32_t *id_row);
int uggpu_loaddata(uggpu_ctx *ctx, const char *filename, int nvd, const int *vec, int64_t nnode, const int32_t *id_level,
                   const int32_t *id_row, uggpu_data_general *general_out /* or NULL */);

/* ---- multigrid cycle, np/procs/iter.cc:7741-7949 Lmgc ------------------------------------------------ */
/* Base solver hook: called with the stream drained when the recursion reaches baselevel.  It must
 * turn the defect b into (correction c, updated defect b) on that level exactly like
 * NP_LINEAR_SOLVER::Solver (np/procs/ls.h:79-132), using uggpu_vec_download/upload.  NULL selects
 * the built-in device solver `ls $I lu` (dense LU without pivoting in vector-index order,
 * np/algebra/ugiter.cc:3657 l_lrdecomp + :4444 l_luiter, iterated like ls.cc:637-749). */
typedef int (*uggpu_base_solver_fn)(void *user, uggpu_ctx *ctx, int level, int c, int b, int A);

typedef struct uggpu_lmgc_cfg {
  int    nu1, nu2, gamma;            /* $n1 $n2 $g     (iter.cc:7637-7642)                     */
  int    baselevel;                  /* $b                                                      */
  double smooth_damp[UGGPU_MAX_BS];  /* jac $damp      (iter.cc:771)                            */
  double cycle_damp[UGGPU_MAX_BS];   /* lmgc $damp     (iter.cc:7665) passed to the prolongation */
  int    t;                          /* vector handle of the temporary np->t (iter.cc:7810)     */
  int    base_maxit;                 /* base `ls $m`                                            */
  double base_reduction;             /* base `ls $red`                                          */
  double base_abslimit;              /* base `ls $abslimit` (default 1e-10, npscan)             */
  uggpu_base_solver_fn base_solver;  /* NULL = built-in device LU                               */
  void  *base_user;
  int    fused;                      /* 0: one kernel per reference call (op-for-op mirror);
                                        1: fused kernels (identical results, fewer passes)      */
  int    smoother;                   /* UGGPU_SM_*: class of the pre- and post-smoother ($S); the fused schedule exists for
                                        jac, the other classes always run one kernel group per reference call */
  int    smoother_L;                 /* ilu: matrix handle that receives the decomposition on every level above the base
                                        level (NP_SMOOTHER.L, allocated by ILUPreProcess iter.cc:5459)                  */
  double ilu_beta[UGGPU_MAX_BS];     /* ilu $beta      (iter.cc:5423)                           */
  int    level_opt;                  /* transfer $L (transfer.cc:574): after the post-smoothing of every level above the base level the
                                        transfer's AdaptCorrection runs (iter.cc:7944 -> transfer.cc:812 -> MinimizeLevel :488); its two
                                        scalars are parallel sums, so with it the cycle agrees with the reference to rounding (1e-12),
                                        not bit for bit; runs the one-kernel-per-call schedule                                     */
} uggpu_lmgc_cfg;

int uggpu_lmgc_preprocess(uggpu_ctx*, const uggpu_lmgc_cfg*, int level, int A);   /* LmgcPreProcess iter.cc:7707 */
/* MinimizeLevel np/procs/transfer.cc:488 (the AdaptCorrection hook of `transfer $L`, :812): t = A c; a0 = (t, b); a1 = |t|^2;
 * c *= 1 + a0/a1; b -= (a0/a1) t -- the correction scaled so that the defect is minimal along it.  t: work vector. */
int uggpu_minimize_level(uggpu_ctx*, int level, int c, int b, int A, int t);
int uggpu_lmgc(uggpu_ctx*, const uggpu_lmgc_cfg*, int level, int c, int b, int A); /* Lmgc          iter.cc:7741 */

/* ---- linear solver, np/procs/ls.cc:562-749 ------------------------------------------------------------ */
typedef struct uggpu_lresult {       /* LRESULT, np/procs/ls.h:72-77 */
  int    error_code;
  int    converged;
  int    number_of_linear_iterations;
  double first_defect[UGGPU_MAX_BS];
  double last_defect[UGGPU_MAX_BS];
} uggpu_lresult;

/* LinearDefect ls.cc:562: b -= A x on levels bl..level, ON_SURFACE */
int uggpu_ls_defect(uggpu_ctx*, int bl, int level, int x, int b, int A);
/* LinearResiduum ls.cc:577: last_defect = dnrm2x(bl..level, ON_SURFACE, b) */
int uggpu_ls_residuum(uggpu_ctx*, int bl, int level, int b, uggpu_lresult *res);
/* LinearSolver ls.cc:637 with Iter = lmgc, Update = LSUpdate (ls.cc:869).  `res->last_defect`
 * must hold the residuum on entry (as in the reference, where Residuum is called first).
 * history (may be NULL) receives last_defect[0..bs) after every iteration: history[it*bs+i]. */
int uggpu_ls_solve(uggpu_ctx*, const uggpu_lmgc_cfg*, int bl, int level, int x, int b, int A, int c,
                   int maxiter, const double *abslimit, const double *reduction,
                   uggpu_lresult *res, double *history);

/* ---- Krylov accelerators around the cycle, np/procs/ls.cc ---------------------------------------------
 * Same conventions as uggpu_ls_solve: res->last_defect holds the residuum on entry, history[it*bs+i].
 * ddotw np/np.h:224, np/algebra/ugblas.cc:3023: sum_i w[i] * (x_i, y_i) over the components */
int uggpu_ddotw(uggpu_ctx*, int fl, int tl, int mode, int x, int y, const double *w /* [bs] */, double *a);
/* class `cg`: LinearSolver ls.cc:637 with CGPrepare :976, CGUpdate :989-1027, CGClose :1159; p, t = the vector handles of
 * NP_CG.p / NP_CG.t (ls.cc:111-124), c = the correction vector of LinearSolver */
int uggpu_cg_solve(uggpu_ctx*, const uggpu_lmgc_cfg*, int bl, int level, int x, int b, int A, int c, int p, int t,
                   int maxiter, const double *abslimit, const double *reduction, uggpu_lresult *res, double *history);
/* class `bcgs`: BCGSSolver ls.cc:1864-2062 with Iter = the cycle and B = NULL; work = the handles of NP_BCGS r p v s t q
 * (ls.cc:165-187), weight = `$weight` as given (squared internally like BCGSInit :1757), restart_every = `$R` (0: never).
 * number_of_linear_iterations counts two per completed pass like the reference; history has one entry per pass. */
int uggpu_bcgs_solve(uggpu_ctx*, const uggpu_lmgc_cfg*, int bl, int level, int x, int b, int A, const int *work /* [6] */,
                     const double *weight /* [bs] */, int restart_every, int maxiter, const double *abslimit,
                     const double *reduction, uggpu_lresult *res, double *history);

/* ---- synthetic hierarchies generated on the device (bench input only; no reference analogue:
 * UG's grid manager needs ~2.5 kB per unknown) -------------------------------------------- */
#define UGGPU_SYNTH_P1_SIMPLEX    0   /* P1 Poisson on Kuhn triangles (nz = 0) / tetrahedra, scalar          */
#define UGGPU_SYNTH_Q1_POISSON    1   /* Q1 Poisson on cubes, scalar, 27-point rows                          */
#define UGGPU_SYNTH_Q1_ELASTICITY 2   /* Q1 linear elasticity on cubes, 3x3 blocks (E = 1, nu = 0.3)         */
#define UGGPU_SYNTH_P1_VARCOEF    3   /* as P1_SIMPLEX with a smoothly VARYING diffusion coefficient per edge: no two rows share
                                         their values, so neither shared value tables nor the stencil kernels apply -- the
                                         explicit (general) path of every kernel, at size                                 */
/* Structured nx*ny*nz cubic cells on level 0 of the unit square (nz = 0) or cube, uniformly refined `top` times.
 * Creates levels 0..top with matrix handle A, the per-row flags of a uniformly refined UG multigrid, Dirichlet
 * identity rows (VECSKIP) on the whole boundary and the standard P/R; sets FULLREFINELEVEL = top. */
int uggpu_synth_hierarchy(uggpu_ctx*, int kind, int nx, int ny, int nz, int top, int A);
/* load vector of the constant right-hand side 1 (0 on Dirichlet rows) into vector `vec` of `level` */
int uggpu_synth_rhs(uggpu_ctx*, int level, int vec);

/* ---- multi-GPU: one context (one process) per GPU, NCCL over NVLink ---------------------------------------------------
 * Replaces, for the hot path, the ppif/DDD call sites of SURVEY.md 2.2: the interface exchange of
 * l_vector_consistent (np/algebra/ugblas.cc:398) becomes a halo copy inside the SpMV-type entry points, the global
 * sums of ddot/dnrm2 (UG_GlobalSumNDOUBLE, parallel/dddif/support.cc:526) an ncclAllReduce inside the reductions,
 * coarse-level agglomeration (np/procs/amgtransfer.cc:246) a replicated coarse hierarchy.  All of it is implicit:
 * the entry points above behave identically on a partitioned hierarchy. */
int uggpu_comm_unique_id(void *out128);                       /* rank 0: ncclGetUniqueId; broadcast the 128 bytes   */
int uggpu_comm_init(uggpu_ctx *ctx, int nranks, int rank, const void *id128);
int uggpu_comm_destroy(uggpu_ctx *ctx);
int uggpu_comm_size(uggpu_ctx *ctx);
int uggpu_comm_rank(uggpu_ctx *ctx);
int64_t uggpu_comm_exchanges(uggpu_ctx *ctx);                 /* halo exchanges issued so far                        */
/* how the halo exchanges of this context travel: 0 no communicator / nothing exchanged yet, 1 ncclSend/ncclRecv,
 * 2 peer-memory windows (CUDA IPC): one push + one wait/unpack kernel per exchange, 3 peer-memory ghost rows: the neighbours'
 * vectors are mapped (CUDA IPC) and interface rows are stored straight into their ghost rows -- by the producing kernel's
 * epilogue in the fused cycle, by one small kernel otherwise -- with one flag wait at the head of the consuming kernel */
#define UGGPU_TRANSPORT_NONE   0
#define UGGPU_TRANSPORT_NCCL   1
#define UGGPU_TRANSPORT_WINDOW 2
#define UGGPU_TRANSPORT_GHOST  3
int uggpu_comm_transport(uggpu_ctx *ctx);
/* ModelP vector consistency, np/algebra/ugblas.cc:398 l_vector_consistent, :1035 l_vector_collect, :740 l_ghostvector_consistent
 * (SURVEY.md 8 a13).  Rows are owned by exactly one rank (owner computes, ghost COLUMNS at the tail of every vector), so the sums
 * over border copies of the first two have nothing to add; the copy of the owners' values into the neighbours' ghost copies is
 * what every entry point does by itself before a kernel reads ghost columns.  The same exchange, explicitly (peer-memory windows
 * over NVLink, or ncclSend/ncclRecv): no-op on one GPU and on levels every rank holds completely. */
int uggpu_l_ghostvector_consistent(uggpu_ctx *ctx, int level, int x);
/* ADDITIVE vectors of a caller that assembles element by element (every rank has added its elements' contributions to all vectors of those
 * elements, also to its ghost rows): l_vector_collect (np/algebra/ugblas.cc:1035) -- the master copy gets the sum over all copies, the
 * ghost rows 0 -- and l_vector_consistent (:398) -- every copy gets the sum.  Ghost segments travel back to their owners (ncclSend/ncclRecv),
 * the owner adds them in the order of its neighbour list (one addition per copy; vectors with more than two copies: "up to summation
 * order", like the reference's interface order).  No-op on one GPU and on levels every rank holds completely. */
int uggpu_l_vector_collect(uggpu_ctx *ctx, int level, int x);
int uggpu_l_vector_consistent(uggpu_ctx *ctx, int level, int x);
/* As uggpu_synth_hierarchy on a px*py*pz rank array (element partition into equal boxes of base cells = RCB of
 * parallel/dddif/lbrcb.cc:250 on a structured grid; sons inherit, lbrcb.cc:376; shared vectors are owned by the lowest
 * rank, priority.cc:200).  This rank generates the rows it owns plus ghost columns.  Levels with at most
 * `replicate_below` rows (and level 0) are held completely by every rank. */
int uggpu_synth_hierarchy_part(uggpu_ctx*, int kind, int nx, int ny, int nz, int top, int A,
                               int px, int py, int pz, int rank, int64_t replicate_below);
/* A partition the CALLER supplies -- what a ModelP application knows from DDD: the vectors this rank is master of
 * (parallel/dddif/priority.cc:200-222) are the level's rows (uggpu_level_create with n = their number), the border / ghost copies it
 * reads become n_ghost ghost rows at the tail of every vector, addressed by column indices >= n in uggpu_mat_set / uggpu_transfer_set.
 * Ghost row recv_off[k] + j is the copy of the j-th row neighbour nb_rank[k] sends; send_idx[send_off[k] .. send_off[k+1]) are the owned
 * rows neighbour k needs, in the order of ITS ghost rows (both sides enumerate an interface in the same order, the role of the sorted
 * interface lists of parallel/ddd/if/ifcreate.cc:155-203).  Call after uggpu_level_create, before any vector of the level exists, on
 * every rank (ranks without a neighbour on the level pass nnb = 0).  Replaces l_vector_consistent / l_ghostvector_consistent
 * (np/algebra/ugblas.cc:398, :740) for the hot path: the entry points exchange what they read.  ug_b200/partition.py derives the
 * lists from a UG hierarchy with the reference's own rules (RCB of the element centres lbrcb.cc:250-330, inheritance :376). */
int uggpu_level_set_partition(uggpu_ctx *ctx, int level, int n_ghost, int64_t n_global, int nnb, const int32_t *nb_rank,
                              const int32_t *send_off /* [nnb+1] */, const int32_t *send_idx, const int32_t *recv_off /* [nnb+1] */);
int64_t uggpu_level_n_global(uggpu_ctx *ctx, int level);      /* rows of the level over all ranks                    */
int uggpu_level_is_partitioned(uggpu_ctx *ctx, int level);
int uggpu_synth_global_ids(uggpu_ctx *ctx, int level, int64_t *ids /* [uggpu_level_n] */);
/* host-only views of the partition arithmetic (no GPU needed; used by the CPU tests of the multi-rank logic) */
int uggpu_part_describe(int dim, int cx, int cy, int cz, int px, int py, int pz, int rank, int32_t *out, int cap);
int uggpu_part_local_index(int dim, int cx, int cy, int cz, int px, int py, int pz, int rank, int x, int y, int z);

#ifdef __cplusplus
}
#endif
#endif /* UGGPU_H */
