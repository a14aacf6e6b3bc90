#!/bin/bash
# bash tools/gpu_sten.sh <tag>: GPU tests, then the default bench with the stencil kernel on / off, Q1 and elasticity
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f achieved %.0f GB/s; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["achieved"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0.05}, d["config"].get("defect"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
run() { name=$1; shift; env "$@" timeout 400 python bench.py --no-cpu --steps 6 --e2e-steps 1 $EXTRA > $out/${tag}_$name.json 2>&1; summ $out/${tag}_$name.json $name; }
run sten A=1
run nosten UGGPU_NO_STENCIL=1
run sten_pf0 UGGPU_PF_DIST=0
run sten_pf14208 UGGPU_PF_DIST=14208
EXTRA="--kind q1" run q1 A=1
