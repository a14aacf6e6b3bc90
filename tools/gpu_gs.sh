#!/bin/bash
# bash tools/gpu_gs.sh <tag>: Gauss-Seidel family tests, then benches with the gs / sgs smoothers
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -k "gs or sor or dropin or ops" > $out/${tag}_pytest.log 2>&1; tail -6 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step"%(d["value"],d["ms_per_step"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0}, d.get("trisolve_finest"), d["config"].get("defect"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
timeout 600 python bench.py --no-cpu --steps 4 --e2e-steps 1 --top 6 --smoother gs > $out/${tag}_gs257.json 2>&1; summ $out/${tag}_gs257.json gs-257
timeout 900 python bench.py --no-cpu --steps 3 --e2e-steps 1 --smoother gs > $out/${tag}_gs513.json 2>&1; summ $out/${tag}_gs513.json gs-513
timeout 900 python bench.py --no-cpu --steps 3 --e2e-steps 1 --smoother sgs > $out/${tag}_sgs513.json 2>&1; summ $out/${tag}_sgs513.json sgs-513
