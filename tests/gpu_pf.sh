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
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if k in ("jac","restrict","interpolate","base")})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
}
KIND=""
run d0 UGGPU_PF_DIST=0
run d7104 UGGPU_PF_DIST=7104
run d9472 UGGPU_PF_DIST=9472
run d11840 UGGPU_PF_DIST=11840
run d14208 UGGPU_PF_DIST=14208
run d9472m3 UGGPU_PF_DIST=9472 UGGPU_PF_MODE=3
KIND="--kind q1"
run q1d9472 UGGPU_PF_DIST=9472
run q1d7104 UGGPU_PF_DIST=7104
KIND="--kind elasticity --top 6"
run eld0 UGGPU_PF_DIST=0
run eld9472 UGGPU_PF_DIST=9472 UGGPU_PF_MAXLINES=1000
run eld4736 UGGPU_PF_DIST=4736 UGGPU_PF_MAXLINES=1000
