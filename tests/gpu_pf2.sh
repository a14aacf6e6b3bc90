#!/bin/bash
tag=$1; out=gpurun_out; mkdir -p $out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu --steps 6 --e2e-steps 1 $KIND > $out/${tag}_$name.log 2>&1
  python - "$out/${tag}_$name.log" "$name" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if k in ("smooth","jac","restrict","interpolate")})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
}
KIND=""
run m7 UGGPU_PF_MODE=7
run m15 UGGPU_PF_MODE=15
run m23 UGGPU_PF_MODE=23
run m31 UGGPU_PF_MODE=31
run m7b UGGPU_PF_MODE=7
run m31b UGGPU_PF_MODE=31
KIND="--kind q1"
run q1m7 UGGPU_PF_MODE=7
run q1m31 UGGPU_PF_MODE=31
