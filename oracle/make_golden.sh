#!/bin/bash
# oracle/make_golden.sh -- TEST INFRASTRUCTURE.  Regenerates tests/golden/*.ugh by running the
# UNMODIFIED reference (oracle/_ref/ugoracle{2,3}, built by oracle/Makefile from /root/reference).
# The reference ships no golden vectors for this path, so these dumps --
# outputs of the reference itself on seeded inputs -- are the pinned fixtures.
set -e
cd "$(dirname "$0")"
make -s -j8 ref
G=../tests/golden
mkdir -p $G
# C1 (BASELINE.json configs[0]) scaled to a committable size: 2D P1, unit square, 4 refinements
./_ref/ugoracle2 --grid tri  --refine 4 --damp 0.8 --cycles 10 --dump $G/c1_tri2d_r4.ugh  --ops --solve > /dev/null
# C2 family: 3D P1 tets, 3 refinements (9^3 = 729 unknowns)
./_ref/ugoracle3 --grid tet  --refine 3 --damp 0.6 --cycles 10 --dump $G/c2_tet3d_r3.ugh  --ops --solve > /dev/null
# C4 family: Q1 hexes, 3x3 blocks (linear elasticity), 2 refinements (125 nodes)
./_ref/ugoracle3 --grid hex  --bs 3 --refine 2 --damp 0.6 --cycles 10 --dump $G/c4_hex3d_bs3_r2.ugh --ops --solve > /dev/null
# C5 family: adaptively refined tets (2 uniform + 2 local refinements): partial levels, classes < 3
./_ref/ugoracle3 --grid tet  --refine 2 --adapt 2 --damp 0.6 --cycles 10 --dump $G/c5_tet3d_adapt.ugh --ops --solve > /dev/null
# 2x2 blocks (plane elasticity): Cramer's rule of SolveSmallBlock (block.cc:119), the b = 2 paths of every kernel; Q1 quads with the
# per-call records + solve + Krylov, P1 triangles with sgs as smoother
./_ref/ugoracle2 --grid quad --bs 2 --refine 3 --damp 0.7 --cycles 8 --dump $G/c4_quad2d_bs2_r3.ugh --ops --solve > /dev/null
./_ref/ugoracle2 --grid tri --bs 2 --refine 3 --damp 0.7 --cycles 6 --lean --smoother sgs --dump $G/sgs_tri2d_bs2_r3.ugh --ops --solve > /dev/null
# quads with scalar unknowns (Q1 in 2D), W-cycle
./_ref/ugoracle2 --grid quad --refine 3 --damp 0.8 --gamma 2 --cycles 6 --dump $G/q1_quad2d_r3_w.ugh --ops --solve > /dev/null
# ---- Gauss-Seidel family as smoother: reference classes gs / sgs / sor inside lmgc.  --lean: hierarchy + the
# smoother records + the solve (every dump above also holds the l_lgs / l_ugs / l_lsor / l_usor records of its hierarchy)
./_ref/ugoracle3 --grid tet --refine 3 --smoother gs  --damp 0.9 --cycles 6 --lean --dump $G/gs_tet3d_r3.ugh      --ops --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --smoother sgs --damp 0.8 --cycles 5 --lean --dump $G/sgs_hex3d_bs3_r2.ugh --ops --solve > /dev/null
./_ref/ugoracle3 --grid tet --refine 2 --adapt 2 --smoother sor --damp 1.1 --cycles 6 --lean --dump $G/sor_tet3d_adapt.ugh --ops --solve > /dev/null
# base level with FREE rows (lmgc $b 2: 125 vectors, 27 of them inside): pins the order of additions of the base LU
./_ref/ugoracle3 --grid tet --refine 3 --baselevel 2 --damp 0.6 --cycles 5 --lean --dump $G/lu_tet3d_r3_bl2.ugh --solve > /dev/null
./_ref/ugoracle2 --grid quad --refine 3 --baselevel 2 --damp 0.8 --cycles 4 --lean --dump $G/lu_quad2d_r3_bl2.ugh --solve > /dev/null
# ... and the block variant of l_lrdecomp / l_luiter (3x3 blocks, 27 free vectors on the base level, fill-in)
./_ref/ugoracle3 --grid hex --bs 3 --refine 3 --baselevel 2 --damp 0.6 --cycles 3 --lean --dump $G/lu_hex3d_bs3_r3_bl2.ugh --solve > /dev/null
# IMAT mode of the transfer (`transfer $M`): RestrictByMatrix / InterpolateCorrectionByMatrix on the stored interpolation
# matrices (created with CreateStandardNodeRestProl), scalar and 3x3 blocks; the transfer records use damping factors != 1
./_ref/ugoracle2 --grid tri --refine 3 --damp 0.8 --cycles 6 --imat --dump $G/imat_tri2d_r3.ugh --ops --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --damp 0.6 --cycles 6 --imat --dump $G/imat_hex3d_bs3_r2.ugh --ops --solve > /dev/null
# ---- ILU smoother: class ilu = l_ilubthdecomp + l_luiter inside lmgc; every dump holds, per level, the decomposed
# matrix values (canonical entry order), one l_luiter, one smoother step, and the solve.  beta != 0 exercises the diagonal modification
./_ref/ugoracle3 --grid tet --refine 3 --smoother ilu --beta 0.25 --damp 0.9 --cycles 6 --lean --dump $G/ilu_tet3d_r3.ugh --ops --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --smoother ilu --beta 0.1 --damp 0.8 --cycles 5 --lean --dump $G/ilu_hex3d_bs3_r2.ugh --ops --solve > /dev/null
./_ref/ugoracle3 --grid tet --refine 2 --adapt 2 --smoother ilu --damp 1.0 --cycles 6 --lean --dump $G/ilu_tet3d_adapt.ugh --ops --solve > /dev/null
./_ref/ugoracle2 --grid quad --refine 3 --smoother ilu --beta 0.5 --damp 1.0 --cycles 4 --lean --dump $G/ilu_quad2d_r3.ugh --ops --solve > /dev/null
# ---- Galerkin coarse-grid operators: AssembleGalerkinByMatrix on the stored interpolation matrices after dmatset(coarse, 0)
# (what `npcheck $G` does), cascaded from the top level down; pattern + values of every coarse level.  Last record group of these dumps.
./_ref/ugoracle3 --grid tet --refine 3 --imat --galerkin --lean --cycles 2 --dump $G/galerkin_tet3d_r3.ugh --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --imat --galerkin --lean --cycles 2 --dump $G/galerkin_hex3d_bs3_r2.ugh --solve > /dev/null
./_ref/ugoracle2 --grid tri --refine 3 --imat --galerkin --lean --cycles 2 --dump $G/galerkin_tri2d_r3.ugh --solve > /dev/null
./_ref/ugoracle3 --grid tet --refine 2 --adapt 2 --imat --galerkin --lean --cycles 2 --dump $G/galerkin_tet3d_adapt.ugh --solve > /dev/null
# ---- element-loop assembly: the reference's LocalAssemble (np/procs/assemble.cc:657) + AssembleDirichletBoundary with the
# element kernel of ug_driver.cc's class `fe` (diffusion with a coefficient per element / linear elasticity, Dirichlet values g(x) != 0):
# elements, coordinates, coefficients, and what the loop leaves on every level (matrix values, rhs, sol, VECSKIP).  Tets, hexes (scalar and
# 3x3 blocks), adaptively refined tets (rows of up to 42 entries), triangles, quads with 2x2 blocks
./_ref/ugoracle3 --grid tet --refine 3 --lean --assemble --dump $G/asm_tet3d_r3.ugh > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --lean --assemble --dump $G/asm_hex3d_bs3_r2.ugh > /dev/null
./_ref/ugoracle3 --grid hex --refine 3 --lean --assemble --dump $G/asm_hex3d_r3.ugh > /dev/null
./_ref/ugoracle3 --grid tet --refine 2 --adapt 2 --lean --assemble --dump $G/asm_tet3d_adapt.ugh > /dev/null
./_ref/ugoracle2 --grid tri --refine 4 --lean --assemble --dump $G/asm_tri2d_r4.ugh > /dev/null
./_ref/ugoracle2 --grid quad --bs 2 --refine 3 --lean --assemble --dump $G/asm_quad2d_bs2_r3.ugh > /dev/null
# ---- savedata / loaddata: files written by the reference's SaveData (np/udm/data_io.cc:650, without a multigrid file, modes bin and
# asc) embedded as byte records, the node-ID order of their bodies, the vectors they were written from
./_ref/ugoracle3 --grid tet --refine 2 --lean --savedata /tmp/ugsd_golden_t --dump $G/savedata_tet3d_r2.ugh > /dev/null
./_ref/ugoracle2 --grid quad --bs 2 --refine 2 --lean --savedata /tmp/ugsd_golden_q --dump $G/savedata_quad2d_bs2_r2.ugh > /dev/null
./_ref/ugoracle3 --grid tet --refine 1 --adapt 2 --lean --savedata /tmp/ugsd_golden_a --dump $G/savedata_tet3d_adapt.ugh > /dev/null
# ---- algebraic levels below level 0 (the AMG side): the reference's own AMG transfer numprocs (np/procs/amgtransfer.cc: selectionAMG with
# Ruge-Stueben coarsening and interpolation, clusterAMG with Vanek's aggregation, selectionAMG with averaging interpolation on a greedy independent set for 3x3 blocks -- not `$C Average`, which re-links the vector list and
# re-sorts the matrix lists of the level it coarsens, amgtools.cc:1330-1440) attached to the
# transfer class with $amg build levels -1, -2, ... with Galerkin matrices under a collapsed level 0; the dumps number the levels from 0 and hold
# the flattened hierarchy (matrices, by-matrix transfer stencils of the algebraic levels) and the solve / Krylov records on all of them
AMG_RS='$strongRel 0.25 $C RugeStueben $I RugeStueben $CM Galerkin $vectLimit 20 $hold'
AMG_VANEK='$strongVanek 0.08 $C VanekNeuss $I Vanek $CM Galerkin $vectLimit 10 $hold'
AMG_AVG='$strongRel 0.25 $C Greedy $I Average $CM Galerkin $vectLimit 10 $hold'
./_ref/ugoracle3 --grid tet --refine 3 --collapse --amg selectionAMG "$AMG_RS" --cycles 5 --dump $G/amg_tet3d_rs.ugh --solve > /dev/null
./_ref/ugoracle2 --grid tri --refine 4 --collapse --refine2 1 --amg clusterAMG "$AMG_VANEK" --cycles 5 --lean --dump $G/amg_tri2d_vanek.ugh --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --collapse --amg selectionAMG "$AMG_AVG" --cycles 4 --lean --dump $G/amg_hex3d_bs3_avg.ugh --solve > /dev/null
./_ref/ugoracle2 --grid quad --refine 4 --collapse --refine2 1 --amg selectionAMG "$AMG_RS" --cycles 5 --lean --dump $G/amg_quad2d_rs.ugh --solve > /dev/null
# aggregation with piecewise constant interpolation in 3D (the smoothed one dereferences a NULL interpolation matrix on this grid in the reference itself)
./_ref/ugoracle3 --grid tet --refine 3 --collapse --amg clusterAMG '$strongVanek 0.08 $C VanekNeuss $I PiecewiseConstant $CM Galerkin $vectLimit 10 $hold' --cycles 4 --lean --dump $G/amg_tet3d_vanek_pc.ugh --solve > /dev/null
# ---- level optimisation (`transfer $L`): AdaptCorrection = MinimizeLevel (np/procs/transfer.cc:812, :488) after the post-smoothing of every level
./_ref/ugoracle3 --grid tet --refine 3 --levelopt --cycles 5 --lean --dump $G/lopt_tet3d_r3.ugh --solve > /dev/null
./_ref/ugoracle3 --grid hex --bs 3 --refine 2 --levelopt --cycles 4 --lean --dump $G/lopt_hex3d_bs3_r2.ugh --solve > /dev/null
ls -la $G
