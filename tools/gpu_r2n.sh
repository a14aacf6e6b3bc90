#!/bin/bash
# N-GPU bench runs without the separate parity script.  bash tools/gpu_r2n.sh <tag> <ngpus> <variants...>   (BENCH_ARGS from the environment)
tag=$1; N=$2; shift 2; out=gpurun_out; mkdir -p $out
run() { name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 $BENCH_ARGS > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  python - $out/${tag}_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "N=%d %s %.3e unk/s %.2f ms/step (events %.2f, kernel sum %.2f) e2e %.3e exch/step %s parity %s"%(d["n_gpus"],d["scaling"],d["value"],d["ms_per_step"],d["config"]["ms_per_step_with_kernel_events"],d["config"]["kernel_sum_ms_per_step"],(d.get("e2e") or {}).get("value",0),d["config"]["halo_exchanges_per_step"],(d.get("mgpu_parity") or {}).get("ok")), "dom %.3f ms"%d["roofline"]["avg_ms"], "bytes %.1f GB"%(d["config"]["device_bytes"]/1e9), "setup %.1fs"%d["config"]["setup_s"])
    for i,pr in enumerate(d.get("per_rank_kernels_ms_per_step") or []): print("    rank",i,pr)
    for k,e in d.items():
        if isinstance(e,dict) and "dominant_kernel" in e: print("   ",k,"%.3e unk/s %.2f ms"%(e["value"],e["ms_per_step"]), e.get("kernels_ms_per_step"))
        elif isinstance(e,dict) and "error" in e and k not in ("spmv",): print("   ",k,"ERROR",e["error"][:300])
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1][:-4]+"err").read()[-2500:])
PY
}
for v in "$@"; do
  case $v in
    fused) run fused UGGPU_HALO_VERBOSE=1;;
    xchg) run xchg UGGPU_NO_FUSED_HALO=1;;
    window) run window UGGPU_HALO=window;;
    nccl) run nccl UGGPU_HALO=nccl;;
  esac
done
