"""Multi-GPU parity check, run as  torchrun --nproc-per-node N tests/mgpu_check.py  (N = 2, 4 or 8 GPUs of one box).

ug_b200.mgpu.parity_check for P1 tetrahedra, Q1 hexahedra and 3x3-block elasticity, in the fused and the
one-kernel-per-call schedule: the partitioned solve must reproduce the single-GPU solve -- x and b BIT FOR BIT, the
defect history to 1e-12.  Prints one line `MGPU-CHECK PASS ...` / `FAIL` per case (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ug_b200 import mgpu  # noqa: E402

CASES = [("p1", 5, 1, True), ("p1", 5, 1, False), ("p1", 5, 0, False), ("q1", 4, 1, True), ("elasticity", 4, 1, True), ("elasticity", 3, 0, False), ("p1var", 4, 1, False)]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok_all = True
    for kind, top, fused, small in CASES:
        r = mgpu.parity_check(rank, world, local, kind=kind, top=int(os.environ.get("MGPU_TOP", top)), fused=fused,
                              replicate_below=int(os.environ.get("MGPU_REPL", "5000")), small_levels=small)
        if rank == 0:
            print(f"MGPU-CHECK {'PASS' if r['ok'] else 'FAIL'} " + json.dumps(r), flush=True)
        ok_all = ok_all and r["ok"]
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
