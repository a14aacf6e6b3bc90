// gpuls_np.h -- the `gpuls` numproc family: drop-in GPU replacements for UG's CPU numprocs on the multigrid
// hot path.  Compiled against the UG headers (-D_2 / -D_3) by the application that links libug; reaches the GPU
// only through the C-ABI of include/uggpu.h, bound at run time with dlopen (no CUDA headers here).
//
//   class name (npcreate $c ...)   abstract base                replaces (reference)
//   iter.gpujac                    NP_ITER   (iter.h:68)        iter.jac          iter.cc:894-942
//   iter.gpugs / gpusgs / gpusor   NP_ITER                      iter.gs / sgs / sor  iter.cc:1003-1090, 1353-1490, 4717-4840
//   transfer.gputransfer           NP_TRANSFER (transfer.h:79)  transfer.transfer transfer.cc:553-897 (standard mode)
//   iter.gpulmgc                   NP_ITER                      iter.lmgc         iter.cc:7613-7980
//   linear_solver.gpuls            NP_LINEAR_SOLVER (ls.h:79)   linear_solver.ls  ls.cc:539-905
//   linear_solver.gpucg / gpubcgs  NP_LINEAR_SOLVER             linear_solver.cg / bcgs  ls.cc:939-1160, 1750-2062
//   assemble.gpufe                 NP_ASSEMBLE (assemble.h:178) the element loop of NP_LOCAL_ASSEMBLE: LocalAssemble assemble.cc:657-706 +
//                                                               NPLocalAssemblePostMatrix :624, with a built-in P1/Q1 element kernel
//                                                               ($P poisson|elasticity $E $nu $f <source per component>)
//
// Same option letters as the CPU classes ($A $x $b $c $damp $S $T $n1 $n2 $g $b $t $m $I $red $abslimit $display),
// plus on gpulmgc: $devbase (solve the base level on the device instead of calling the BaseSolver numproc) and
// $unfused (one kernel per reference call).  Usage: call InitGpuLS() once after InitUg(), then e.g.
//   npcreate smooth $c gpujac; npcreate transfer $c gputransfer; npcreate lmgc $c gpulmgc; npcreate mgs $c gpuls;
#ifndef GPULS_NP_H
#define GPULS_NP_H

#include "gm.h"
#include "np.h"

namespace gpuls {
// dlopen()s libuggpu.so (path may be NULL: $UGGPU_LIB, then "libuggpu.so" on the loader path). 0 = ok.
int LoadDeviceLibrary(const char *path);
const char *LastLoadError();
// assemble.gpufe: the application's data for the built-in element kernel -- one coefficient per element (NULL: 1) and the Dirichlet
// value of component `comp` at a boundary vertex (NULL: 0).  In UG both live inside the application's AssembleLocal.
typedef double (*ElemCoefFn)(NS_DIM_PREFIX ELEMENT *e);
typedef double (*DirichletFn)(const double *pos, int comp);
void SetFEData(ElemCoefFn coef, DirichletFn dirichlet);
// savedata / loaddata for DEVICE vectors (reference: SaveData np/udm/data_io.cc:650 with save_without_mg, LoadData :408; the commands
// `savedata` / `loaddata` of ui/commands.cc call those).  Valid inside a PreProcess/PostProcess bracket of a gpuls numproc, i.e. while the
// device mirror of `mg` holds the vectors: the values are taken from / written to the device copies, the VVALUEs are not touched.  Files
// are byte-identical to the reference's.  type: "asc" | "bin"; number = -1: no time step suffix.  0 = ok.
int SaveData(NS_DIM_PREFIX MULTIGRID *mg, const char *name, const char *type, int number, double time, double dt, double ndt, int n,
             NS_DIM_PREFIX VECDATA_DESC **vds);
int LoadData(NS_DIM_PREFIX MULTIGRID *mg, const char *name, const char *type, int number, int n, NS_DIM_PREFIX VECDATA_DESC **vds);
}

START_UGDIM_NAMESPACE
// registers the classes (np/udm/numproc.h:108 CreateClass); 0 = ok
INT InitGpuLS(void);
END_UGDIM_NAMESPACE

#endif
