#!/bin/bash
# bash tools/gpu_variants.sh <tag> <variant>...: default bench with each build variant ug_b200/lib/libuggpu_<variant>.so ("base" = libuggpu.so)
tag=$1; shift; out=gpurun_out; mkdir -p $out
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0.05})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for v in "$@"; do
  lib=$PWD/ug_b200/lib/libuggpu_$v.so; [ "$v" = base ] && lib=$PWD/ug_b200/lib/libuggpu.so
  UGGPU_LIB=$lib timeout 300 python bench.py --no-cpu --steps 3 --e2e-steps 1 $EXTRA > $out/${tag}_$v.json 2>&1; summ $out/${tag}_$v.json $v
done
