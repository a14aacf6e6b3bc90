#!/bin/bash
tag=$1; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -8 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f e2e %.3e"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()}, d.get("trisolve_finest"), d["config"].get("defect"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
timeout 600 python bench.py --no-cpu --steps 10 --e2e-steps 2 > $out/${tag}_p1.json 2>&1; summ $out/${tag}_p1.json p1
timeout 600 python bench.py --no-cpu --steps 10 --e2e-steps 1 --cells 8 --top 5 > $out/${tag}_base729.json 2>&1; summ $out/${tag}_base729.json base729-257
UGGPU_LU_INDEX_ORDER=1 timeout 600 python bench.py --no-cpu --steps 10 --e2e-steps 1 --cells 8 --top 5 > $out/${tag}_base729_idx.json 2>&1; summ $out/${tag}_base729_idx.json base729-257-indexorder
