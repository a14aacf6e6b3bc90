#!/bin/bash
# multi-GPU round: parity test, then bench with the peer-memory and the NCCL halo exchange.  bash tools/gpu_mgpu.sh <tag> <ngpus>
tag=$1; N=$2; out=gpurun_out; mkdir -p $out
UGGPU_HALO_VERBOSE=1 timeout 600 python -m pytest tests/test_mgpu.py -m gpu -x -q 2>&1 | tail -5
run() { name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  grep -h "uggpu: halo" $out/${tag}_$name.err | head -1
  python - $out/${tag}_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%d %.3e unk/s %.2f ms/step e2e %.3e exch %s"%(d["n_gpus"],d["value"],d["ms_per_step"],d["e2e"]["value"],d["config"]["halo_exchanges_total"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1][:-4]+"err").read()[-1500:])
PY
}
run p2p UGGPU_HALO_VERBOSE=1
run nccl UGGPU_HALO=nccl UGGPU_HALO_VERBOSE=1
