#!/bin/bash
# A/B of the software-prefetch distance / mode of the thread-per-row kernels
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 5 --e2e-steps 1 $KIND > $out/${tag}_$name.log 2>&1
  python - "$out/${tag}_$name.log" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"]))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
}
KIND=""
run d0 UGGPU_PF_DIST=0
run d2368 UGGPU_PF_DIST=2368
run d4736 UGGPU_PF_DIST=4736
run d9472 UGGPU_PF_DIST=9472
run d18944 UGGPU_PF_DIST=18944
run d9472m1 UGGPU_PF_DIST=9472 UGGPU_PF_MODE=1
run d9472m7 UGGPU_PF_DIST=9472 UGGPU_PF_MODE=7
KIND="--kind q1"
run q1d0 UGGPU_PF_DIST=0
run q1d9472 UGGPU_PF_DIST=9472
KIND="--kind elasticity --top 6"
run eld0 UGGPU_PF_DIST=0
run eld9472 UGGPU_PF_DIST=9472
