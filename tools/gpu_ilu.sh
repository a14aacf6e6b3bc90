#!/bin/bash
# bash tools/gpu_ilu.sh <tag>: benches with the ilu smoother (257^3 and 513^3)
tag=$1; out=gpurun_out; mkdir -p $out
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step"%(d["value"],d["ms_per_step"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0}, d.get("trisolve_finest"), d["config"].get("defect"), "preprocess_s", d["config"].get("preprocess_s"), "bytes", d["config"].get("device_bytes"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
timeout 600 python bench.py --no-cpu --steps 4 --e2e-steps 1 --top 6 --smoother ilu > $out/${tag}_ilu257.json 2>&1; summ $out/${tag}_ilu257.json ilu-257
timeout 900 python bench.py --no-cpu --steps 3 --e2e-steps 1 --smoother ilu > $out/${tag}_ilu513.json 2>&1; summ $out/${tag}_ilu513.json ilu-513
