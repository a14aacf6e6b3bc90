#!/bin/bash
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synth.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["frac"],r["cycle_frac"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if k in ("smooth","restrict","interpolate","jac","base")})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for v in "k1 UGGPU_TR_K_INTERP=1 UGGPU_TR_K_RESTRICT=1" "k2 UGGPU_TR_K_INTERP=2 UGGPU_TR_K_RESTRICT=2" "k4 UGGPU_TR_K_INTERP=4 UGGPU_TR_K_RESTRICT=4" "k1b UGGPU_TR_K_INTERP=1 UGGPU_TR_K_RESTRICT=1" "k2early UGGPU_TR_K_INTERP=2 UGGPU_TR_K_RESTRICT=2 UGGPU_PF_MODE=127"; do
  set -- $v; name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu --steps 6 --e2e-steps 1 > $out/${tag}_ab_$name.json 2>&1; summ $out/${tag}_ab_$name.json $name
done
