"""Drop-in test of the `gpuls` numproc family inside the UNMODIFIED reference: oracle/_ref/ugoracle{2,3} link libug
(compiled from /root/reference) with ug_b200/host/gpuls_np.cc, register iter.gpujac / transfer.gputransfer /
iter.gpulmgc / linear_solver.gpuls via InitGpuLS(), run the SAME numproc script once with the CPU classes and once
with every GPU/CPU mix, and compare the VVALUEs UG holds afterwards (bit-exact when the base solver is the CPU numproc,
1e-12 with the device LU) and LRESULT.  The binaries are built where /root/reference exists (oracle/Makefile) and
travel to the GPU box; the test is skipped if they are absent."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ug_b200", "lib", "libuggpu.so")

AMG_RS = "$strongRel 0.25 $C RugeStueben $I RugeStueben $CM Galerkin $vectLimit 20"
AMG_VANEK = "$strongVanek 0.08 $C VanekNeuss $I Vanek $CM Galerkin $vectLimit 10"
AMG_VANEK_PC = "$strongVanek 0.08 $C VanekNeuss $I PiecewiseConstant $CM Galerkin $vectLimit 60"
# averaging interpolation (the one the reference offers for systems) on a greedy independent set.  Not `$C Average`: CoarsenAverage re-links
# the vector list of the level it coarsens and re-sorts its matrix lists (np/algebra/amgtools.cc:1330-1440), so two successive solves of the
# reference itself run on differently ordered levels and differ in the last bits -- nothing a run-after-run comparison can be pinned on
AMG_AVG = "$strongRel 0.25 $C Greedy $I Average $CM Galerkin"
CASES = [
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle2", ["--grid", "tri", "--refine", "5", "--damp", "0.8", "--cycles", "6"]),
    ("ugoracle2", ["--grid", "quad", "--refine", "3", "--damp", "0.8", "--gamma", "2", "--cycles", "5"]),
    # Gauss-Seidel family as smoother (iter.gpugs / gpusgs / gpusor against the reference's gs / sgs / sor)
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--smoother", "gs", "--damp", "0.9", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--smoother", "sgs", "--damp", "0.8", "--cycles", "4"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--smoother", "sor", "--damp", "1.1", "--cycles", "5"]),
    # base level with free rows (lmgc $b 2)
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--baselevel", "2", "--damp", "0.6", "--cycles", "5"]),
    # transfer $M: stored interpolation matrices (RestrictByMatrix / InterpolateCorrectionByMatrix) against gputransfer $M
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--imat", "--damp", "0.6", "--cycles", "5"]),
    # ILU smoother (iter.gpuilu against the reference's ilu: l_ilubthdecomp with $beta + l_luiter)
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--smoother", "ilu", "--beta", "0.25", "--damp", "0.9", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--smoother", "ilu", "--beta", "0.1", "--damp", "0.8", "--cycles", "4"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--smoother", "ilu", "--damp", "1.0", "--cycles", "5"]),
    # 2x2 blocks (plane elasticity)
    ("ugoracle2", ["--grid", "quad", "--bs", "2", "--refine", "4", "--damp", "0.7", "--cycles", "6"]),
    # ---- at size: the storage forms that only switch on for large levels (fixed-width layout, L2 prefetch, multi-block reductions,
    # the stencil-rows / exception-rows kernels, the row-class transfer) against the REFERENCE ITSELF, not only against the port.
    # --nokrylov: the four ls / lmgc mixes only (the reference's own cg / bcgs runs at this size take minutes of CPU time)
    ("ugoracle3", ["--grid", "tet", "--refine", "5", "--damp", "0.6", "--cycles", "5", "--nokrylov"]),                 # 33^3 = 35 937 unknowns
    ("ugoracle3", ["--grid", "tet", "--refine", "6", "--damp", "0.6", "--cycles", "4", "--nokrylov"]),                 # 65^3 = 274 625 unknowns
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "4", "--damp", "0.6", "--cycles", "4", "--nokrylov"]),      # 17^3 nodes x 3
    ("ugoracle3", ["--grid", "hex", "--refine", "5", "--damp", "0.6", "--cycles", "4", "--nokrylov"]),                 # Q1 Poisson 33^3: 27-point rows
    ("ugoracle2", ["--grid", "tri", "--refine", "6", "--damp", "0.8", "--cycles", "6"]),                               # C1 verbatim: 65^2 = 4 225 unknowns, 7 levels
    ("ugoracle2", ["--grid", "tri", "--refine", "9", "--damp", "0.8", "--cycles", "4", "--nokrylov"]),                 # 513^2 = 263 169 unknowns
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--adapt", "2", "--damp", "0.6", "--cycles", "4", "--nokrylov"]),   # adaptive on 17^3
    # ---- assemble.gpufe against the reference's NP_LOCAL_ASSEMBLE loop (np/procs/assemble.cc:657) with the same element
    # kernel: MVALUEs, rhs, Dirichlet values and VECSKIP of every level bit for bit, then gpuls inside gpufe's bracket vs the CPU solve
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--damp", "0.6", "--cycles", "5", "--nokrylov", "--assemble"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--damp", "0.6", "--cycles", "4", "--nokrylov", "--assemble"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--damp", "0.6", "--cycles", "4", "--nokrylov", "--assemble"]),
    ("ugoracle2", ["--grid", "quad", "--bs", "2", "--refine", "3", "--damp", "0.7", "--cycles", "4", "--nokrylov", "--assemble"]),
    ("ugoracle3", ["--grid", "hex", "--refine", "4", "--damp", "0.6", "--cycles", "4", "--nokrylov", "--assemble"]),     # Q1 17^3
    ("ugoracle3", ["--grid", "tet", "--refine", "5", "--damp", "0.6", "--cycles", "4", "--nokrylov", "--assemble"]),     # 33^3
    # ---- transfer $L / gputransfer $L: level optimisation (MinimizeLevel, np/procs/transfer.cc:488) inside the cycle; agreement to rounding
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--levelopt", "--damp", "0.6", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--levelopt", "--damp", "0.6", "--cycles", "4"]),
    # transfer $D (AssembleDirichletBoundary in the transfer's PreProcess) and the nested-iteration hooks InterpolateNewVectors / ProjectSolution
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--transferD", "--hooks", "--damp", "0.6", "--cycles", "4", "--nokrylov"]),
    # ---- algebraic levels: `transfer $amg amgt` / `gputransfer $amg amgt` with the reference's own AMG transfer numproc
    # (np/procs/amgtransfer.cc) building levels -1, -2, ... below a collapsed level 0 in every PreProcess; the device cycle runs on them
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--collapse", "--cycles", "5", "--amg", "selectionAMG", AMG_RS]),                          # Ruge-Stueben
    ("ugoracle2", ["--grid", "tri", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--amg", "clusterAMG", AMG_VANEK]),       # Vanek aggregation, one geometric level above
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--collapse", "--cycles", "4", "--amg", "selectionAMG", AMG_AVG + " $vectLimit 10"]),   # 3x3 blocks, averaging
    ("ugoracle3", ["--grid", "tet", "--refine", "5", "--collapse", "--refine2", "1", "--cycles", "4", "--nokrylov", "--amg", "selectionAMG", AMG_AVG + " $vectLimit 40"]),   # 65^3 on 33^3 on 4 algebraic levels
    # ---- gputransfer $gpuamg: the algebraic levels are built by the device library itself (uggpu_amg_coarsen_rs / _vanek) and exist on the device
    # only; the reference side runs its own AMG numproc -- same levels, same bits.  Device-resident solve with the device base solver only
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--collapse", "--cycles", "5", "--nokrylov", "--amg", "selectionAMG", AMG_RS, "--gpuamg", "RugeStueben $theta 0.25 $vectLimit 20"]),
    ("ugoracle2", ["--grid", "quad", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--nokrylov", "--amg", "selectionAMG", AMG_RS, "--gpuamg", "RugeStueben $vectLimit 20"]),
    ("ugoracle2", ["--grid", "tri", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--nokrylov", "--amg", "clusterAMG", AMG_VANEK, "--gpuamg", "Vanek $theta 0.08 $vectLimit 10"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "4", "--nokrylov", "--amg", "clusterAMG", AMG_VANEK_PC, "--gpuamg", "VanekPC $theta 0.08 $vectLimit 60"]),
]
IDS = ["tet-r3", "tet-adaptive", "hex-bs3", "tri-r5", "quad-W", "tet-gs", "hex-bs3-sgs", "tet-adaptive-sor", "tet-baselevel2", "hex-bs3-imat",
       "tet-ilu-beta", "hex-bs3-ilu-beta", "tet-adaptive-ilu", "quad-bs2", "tet-r5-33^3", "tet-r6-65^3", "hex-bs3-r4", "hex-q1-r5-33^3", "tri-r6-C1", "tri-r9-513^2",
       "tet-r4-adaptive", "assemble-tet-r3", "assemble-hex-bs3", "assemble-tet-adaptive", "assemble-quad-bs2", "assemble-hex-q1-r4", "assemble-tet-r5-33^3",
       "tet-r4-levelopt", "hex-bs3-levelopt", "tet-adaptive-transferD-hooks",
       "amg-tet-ruge-stueben", "amg-tri-vanek-refine2", "amg-hex-bs3-greedy-average", "amg-tet-65^3-on-33^3-greedy-average",
       "gpuamg-tet-ruge-stueben", "gpuamg-quad-ruge-stueben-refine2", "gpuamg-tri-vanek-refine2", "gpuamg-tet-33^3-on-17^3-vanek-pc"]


@pytest.mark.parametrize("exe,args", CASES, ids=IDS)
def test_gpuls_numprocs_inside_ug(exe, args):
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    out = subprocess.run([path] + args + ["--gpu", LIB], capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith(("PASS", "FAIL", "gpuls"))]
    assert out.returncode == 0, "\n".join(lines) + out.stderr[-2000:]
    want = (4 if "--nokrylov" in args else 6) + (6 if "--assemble" in args else 0) + (1 if "--hooks" in args else 0)
    if "--gpuamg" in args:
        want = 1        # the device-resident solve with the device base solver: the only mix with algebraic levels that exist on the device alone      # 4 ls/lmgc mixes [+ gpucg + gpubcgs] [+ gpufe, gpuls inside its bracket, savedata / loaddata bin + asc]
    assert sum(l.startswith("PASS") for l in lines) == want, lines
    assert lines[-1] == "gpuls drop-in: 0 failure(s)"
