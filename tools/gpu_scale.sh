#!/bin/bash
# bash tools/gpu_scale.sh <tag> "<N ...>": multi-GPU parity test, then the weak-scaling bench at each N (peer-memory halo exchange)
tag=$1; out=gpurun_out; mkdir -p $out
UGGPU_HALO_VERBOSE=1 timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q 2>&1 | tail -4
for N in $2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_n$N.json 2> $out/${tag}_n$N.err
  python - $out/${tag}_n$N.json $N <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N=%d %.3e unk/s %.2f ms/step e2e %.3e exch %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"]["halo_exchanges_total"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()})
except Exception as e:
    print("N=%s failed"%sys.argv[2], e); print(open(sys.argv[1][:-4]+"err").read()[-1500:])
PY
done
