#!/bin/bash
# bash tools/gpu_overlap.sh <tag> <N> [cells]: multi-GPU parity, then the bench with and without the interior/interface overlap
tag=$1; N=$2; cells=${3:-4}; out=gpurun_out; mkdir -p $out
UGGPU_HALO_VERBOSE=1 timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q 2>&1 | tail -4
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%d %.3e unk/s %.3f ms/step e2e %.3e exch %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"]["halo_exchanges_total"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1][:-4]+"err").read()[-1500:])
PY
}
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 --cells $cells > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  show $out/${tag}_$name.json $name
}
run overlap UGGPU_X=1
run serial UGGPU_NO_OVERLAP=1
run overlap2 UGGPU_X=1
