#!/bin/bash
# bash tools/gpu_ab2.sh <tag> "<name ENV=..>" ...: full GPU tests, then one default bench per variant
tag=$1; shift; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu --steps 10 --e2e-steps 1 > $out/${tag}_$name.json 2>&1
  python - $out/${tag}_$name.json $name <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
done
