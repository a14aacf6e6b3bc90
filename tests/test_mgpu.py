"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): tests/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_solve_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", os.path.join(ROOT, "tests", "mgpu_check.py")], capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("MGPU-CHECK")]
    assert out.returncode == 0 and len(lines) == 7 and all("PASS" in l for l in lines), out.stdout[-3000:] + out.stderr[-3000:]


def test_caller_supplied_partition_of_a_ug_hierarchy():
    """uggpu_level_set_partition: a UG golden hierarchy partitioned by the reference's RCB / ownership rules (ug_b200/partition.py) solves on
    2 or 4 GPUs to the bits the reference left in its VECTORs (tests/part_check.py)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n >= 4 else 2
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29519", os.path.join(ROOT, "tests", "part_check.py")], capture_output=True, text=True, timeout=900)
    lines = [l for l in out.stdout.splitlines() if l.startswith("PART-CHECK")]
    assert out.returncode == 0 and len(lines) == 5 and all("PASS" in l for l in lines), out.stdout[-3000:] + out.stderr[-3000:]
