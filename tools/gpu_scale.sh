#!/bin/bash
# bash tools/gpu_scale.sh <tag> "<N ...>" [strongN]: multi-GPU parity test, the weak-scaling bench at each N (peer-memory halo
# exchange), and optionally the STRONG-scaling point (the 513^3 problem of one GPU split over strongN GPUs: --cells 4/P per direction)
tag=$1; out=gpurun_out; mkdir -p $out
UGGPU_HALO_VERBOSE=1 timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%d %.3e unk/s %.2f ms/step e2e %.3e exch %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"]["halo_exchanges_total"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1][:-4]+"err").read()[-1500:])
PY
}
for N in $2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_n$N.json 2> $out/${tag}_n$N.err
  show $out/${tag}_n$N.json weak
done
if [ -n "$3" ]; then
  N=$3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 --cells 2 > $out/${tag}_strong$N.json 2> $out/${tag}_strong$N.err
  show $out/${tag}_strong$N.json strong
fi
