#!/bin/bash
# bash tools/gpu_n2.sh <tag> <ngpus>: multi-GPU tests (caller-supplied partitions against the reference's dumps) + the driver's bench command at N GPUs
tag=$1; N=$2; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_mgpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_n$N.json 2> $out/${tag}_n$N.err
python - $out/${tag}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d %s %.3e unk/s %.3f ms/step kernel sum %.3f" % (d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["config"]["kernel_sum_ms_per_step"]), "parity", {k: d["mgpu_parity"][k] for k in ("ok", "x_bitexact", "b_bitexact", "hist_relerr")})
    for k in ("weak", "q1_poisson_strong", "elasticity_3x3_strong", "varying_coefficient_strong"):
        if k in d: print("  ", k, d[k].get("ms_per_step"), d[k].get("value"), d[k].get("error"))
except Exception as e:
    print("failed", e); print(open(sys.argv[1][:-4] + "err").read()[-1500:])
PY
