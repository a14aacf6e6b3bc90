#!/bin/bash
# generic A/B: bash tools/gpu_ab.sh <tag> "<name> ENV=.. ENV=.." ...   (runs pytest first, then one p1 bench per variant, then q1 for each)
tag=$1; shift; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for spec in "$@"; do
  set -- $spec; name=$1; shift
  for kind in p1 q1; do
    env "$@" timeout 300 python bench.py --no-cpu --steps 6 --e2e-steps 1 --kind $kind > $out/${tag}_${name}_$kind.log 2>&1
    python - "$out/${tag}_${name}_$kind.log" "$name/$kind" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if k in ("smooth","jac","restrict","interpolate")})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-800:])
PY
  done
done
