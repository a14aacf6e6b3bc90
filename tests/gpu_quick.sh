#!/bin/bash
# parity tests under a timeout (a hung bulk copy must not hang the box), then bench lines.  bash tests/gpu_quick.sh <tag> [bench args...]
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/${tag}_pytest.log
cat $out/${tag}_pytest.log | tail -12
if grep -q "failed\|error\|Killed\|Terminated" $out/${tag}_pytest.log; then echo "TESTS FAILED - skipping bench"; exit 1; fi
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f; e2e %.3e"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"]))
    print("   ", {k:(v["launches"],round(v["ms"],2)) for k,v in d["kernels"].items()})
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for kind in p1 q1; do
  timeout 600 python bench.py --no-cpu --steps 5 --e2e-steps 1 --kind $kind "$@" > $out/${tag}_bench_${kind}.log 2>&1; summ $out/${tag}_bench_${kind}.log $kind
done
UGGPU_NO_TMA=1 timeout 600 python bench.py --no-cpu --steps 5 --e2e-steps 1 "$@" > $out/${tag}_bench_p1_notma.log 2>&1; summ $out/${tag}_bench_p1_notma.log p1-notma
