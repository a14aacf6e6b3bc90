"""Host logic of the `gpuls` numproc family on a machine WITHOUT a GPU: the drop-in cases of tests/test_dropin.py inside the unmodified
reference (oracle/_ref/ugoracle{2,3}), with tests/standin/libuggpu_standin.so in place of libuggpu.so.  The stand-in answers the C-ABI
calls the numprocs make with the oracle's plain-C restatement, so what is exercised here is everything ABOVE the C-ABI: flattening of
UG's lists, level numbering (algebraic levels below 0 -> device levels), upload caching per PreProcess bracket, the base-solver
callback, LRESULT.  Results must equal the CPU numprocs' bit for bit (the restatement is bit-exact).  The CUDA library is tested by
the same cases in tests/test_dropin.py (`-m gpu`)."""
import os
import subprocess

import pytest

from test_dropin import AMG_AVG, AMG_RS, AMG_VANEK, AMG_VANEK_PC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "standin", "uggpu_standin.cc")
LIB = os.path.join(ROOT, "tests", "standin", "libuggpu_standin.so")


@pytest.fixture(scope="module")
def standin():
    deps = [SRC, os.path.join(ROOT, "oracle", "ugport.c"), os.path.join(ROOT, "oracle", "ugport.h"), os.path.join(ROOT, "include", "uggpu.h"),
            os.path.join(ROOT, "ug_b200", "csrc", "amg_host.inc")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        obj = LIB[:-3] + ".ugport.o"
        inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle")]
        subprocess.run(["gcc", "-O1", "-ffp-contract=off", "-fPIC", "-c", deps[1], "-o", obj] + inc, check=True)
        subprocess.run(["g++", "-std=c++14", "-O1", "-ffp-contract=off", "-fPIC", "-shared", SRC, obj, "-o", LIB, "-lm", "-ldl"] + inc, check=True)
    return LIB


CASES = [
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--damp", "0.6", "--cycles", "6"]),
    ("ugoracle2", ["--grid", "tri", "--refine", "5", "--damp", "0.8", "--cycles", "6"]),
    ("ugoracle2", ["--grid", "quad", "--refine", "3", "--damp", "0.8", "--gamma", "2", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--smoother", "gs", "--damp", "0.9", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--smoother", "sgs", "--damp", "0.8", "--cycles", "4"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--smoother", "sor", "--damp", "1.1", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--baselevel", "2", "--damp", "0.6", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--imat", "--damp", "0.6", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--smoother", "ilu", "--beta", "0.25", "--damp", "0.9", "--cycles", "5"]),
    ("ugoracle2", ["--grid", "quad", "--bs", "2", "--refine", "4", "--damp", "0.7", "--cycles", "6"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--levelopt", "--damp", "0.6", "--cycles", "5"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--levelopt", "--damp", "0.6", "--cycles", "4"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--transferD", "--hooks", "--damp", "0.6", "--cycles", "4"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--imat", "--hooks", "--damp", "0.6", "--cycles", "4"]),
    # assemble.gpufe + gpuls inside its bracket + savedata / loaddata of the mirror's vectors (the stand-in borrows the product's host-only format functions)
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--damp", "0.6", "--cycles", "5", "--assemble"]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--damp", "0.6", "--cycles", "4", "--assemble"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "2", "--adapt", "2", "--damp", "0.6", "--cycles", "4", "--assemble"]),
    ("ugoracle2", ["--grid", "quad", "--bs", "2", "--refine", "3", "--damp", "0.7", "--cycles", "4", "--assemble"]),
    # algebraic levels below level 0: `gputransfer $amg amgt` calls the reference's AMG numproc, mirrors levels -1, -2, ... as device levels
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--collapse", "--cycles", "5", "--amg", "selectionAMG", AMG_RS]),
    ("ugoracle2", ["--grid", "tri", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--amg", "clusterAMG", AMG_VANEK]),
    ("ugoracle3", ["--grid", "hex", "--bs", "3", "--refine", "2", "--collapse", "--cycles", "4", "--amg", "selectionAMG", AMG_AVG + " $vectLimit 10"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "4", "--amg", "selectionAMG", AMG_AVG + " $vectLimit 40"]),
    # ---- gputransfer $gpuamg: the algebraic levels are built by the device library itself (uggpu_amg_coarsen_rs / _vanek) and exist on the device
    # only; the reference side runs its own AMG numproc -- same levels, same bits.  Device-resident solve with the device base solver only
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--collapse", "--cycles", "5", "--amg", "selectionAMG", AMG_RS, "--gpuamg", "RugeStueben $theta 0.25 $vectLimit 20"]),
    ("ugoracle2", ["--grid", "quad", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--amg", "selectionAMG", AMG_RS, "--gpuamg", "RugeStueben $vectLimit 20"]),
    ("ugoracle2", ["--grid", "tri", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "5", "--amg", "clusterAMG", AMG_VANEK, "--gpuamg", "Vanek $theta 0.08 $vectLimit 10"]),
    ("ugoracle3", ["--grid", "tet", "--refine", "4", "--collapse", "--refine2", "1", "--cycles", "4", "--amg", "clusterAMG", AMG_VANEK_PC, "--gpuamg", "VanekPC $theta 0.08 $vectLimit 60"]),
] + [
    # algebraic levels with the other smoother classes, the W-cycle and the level optimisation (host logic only: the CUDA kernels of these
    # combinations run on algebraic levels in the GPU suite through the golden dumps with the Jacobi smoother)
    ("ugoracle3", ["--grid", "tet", "--refine", "3", "--collapse", "--refine2", "1", "--cycles", "4", "--amg", "selectionAMG", AMG_RS] + extra)
    for extra in (["--smoother", "gs", "--damp", "0.9"], ["--smoother", "ilu", "--beta", "0.25", "--damp", "0.9"], ["--levelopt"], ["--smoother", "sgs", "--damp", "0.8", "--gamma", "2"])
] + [
    # the stopping criteria of the coarsening loop (amgtransfer.cc:806-826, :1000-1012), the same on both sides: same number of levels, same bits
    ("ugoracle2", ["--grid", "tri", "--refine", "5", "--collapse", "--cycles", "4", "--amg", "selectionAMG", "$strongRel 0.25 $C RugeStueben $I RugeStueben $CM Galerkin " + crit,
                   "--gpuamg", "RugeStueben " + crit])
    for crit in ("$vRedLimit 0.3", "$bandLimit 12", "$mRedLimit 0.6", "$matLimit 3000", "$levelLimit -2")
]
IDS = ["tet-r3", "tet-adaptive", "hex-bs3", "tri-r5", "quad-W", "tet-gs", "hex-bs3-sgs", "tet-adaptive-sor", "tet-baselevel2", "hex-bs3-imat", "tet-ilu-beta",
       "quad-bs2", "tet-levelopt", "hex-bs3-levelopt", "tet-adaptive-transferD-hooks", "hex-bs3-imat-hooks", "assemble-tet-r3", "assemble-hex-bs3", "assemble-tet-adaptive", "assemble-quad-bs2", "amg-tet-ruge-stueben", "amg-tri-vanek-refine2", "amg-hex-bs3-greedy-average", "amg-tet-33^3-on-17^3-greedy-average",
       "gpuamg-tet-ruge-stueben", "gpuamg-quad-ruge-stueben-refine2", "gpuamg-tri-vanek-refine2", "gpuamg-tet-33^3-on-17^3-vanek-pc",
       "amg-gs", "amg-ilu", "amg-levelopt", "amg-sgs-W",
       "gpuamg-vRedLimit", "gpuamg-bandLimit", "gpuamg-mRedLimit", "gpuamg-matLimit", "gpuamg-levelLimit"]


@pytest.mark.parametrize("exe,args", CASES, ids=IDS)
def test_host_numprocs_against_standin(standin, exe, args):
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    # gpucg / gpubcgs around the cycle as well (the stand-in's are the restatement's: bit for bit), except where the reference's own Krylov runs take long
    nokry = ["--nokrylov"] if ("--gpuamg" in args or "--assemble" in args or ("--refine", "4") == tuple(args[2:4]) and "--collapse" in args) else []
    out = subprocess.run([path] + args + nokry + ["--gpu", standin], capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith(("PASS", "FAIL", "gpuls"))]
    assert out.returncode == 0, "\n".join(lines) + out.stderr[-2000:]
    assert sum(l.startswith("PASS") for l in lines) == (1 if "--gpuamg" in args else 4 + (0 if nokry else 2) + (1 if "--hooks" in args else 0) + (6 if "--assemble" in args else 0)), lines
    # bit for bit, also with the "device" base solver (the stand-in's is the restatement of the reference's ls + lu)
    assert all("relerr x=0.000e+00 b=0.000e+00" in l for l in lines if l.startswith("PASS") and "relerr x=" in l and "gpufe bracket" not in l and " vs " not in l), lines
    assert lines[-1] == "gpuls drop-in: 0 failure(s)"


def test_failed_preprocess_leaves_a_clean_state(standin):
    """A PreProcess that fails inside the bracket ($gpuamg on a block system: scalar equations only) must fail loudly and cleanly -- message,
    non-zero result, no crash, the mirror dropped."""
    path = os.path.join(ROOT, "oracle", "_ref", "ugoracle3")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    args = ["--grid", "hex", "--bs", "3", "--refine", "2", "--collapse", "--cycles", "3", "--nokrylov",
            "--amg", "selectionAMG", AMG_AVG + " $vectLimit 10", "--gpuamg", "RugeStueben"]
    out = subprocess.run([path] + args + ["--gpu", standin], capture_output=True, text=True, timeout=300)
    assert out.returncode == 10, (out.returncode, out.stdout[-500:], out.stderr[-500:])            # the driver's "failures" exit code, not a signal
    assert "gputransfer: $gpuamg handles scalar equations" in out.stdout
    assert "FAIL gpuls+gpulmgc, device base solver: PreProcess" in out.stdout
    assert out.stdout.strip().splitlines()[-1] == "gpuls drop-in: 1 failure(s)"
