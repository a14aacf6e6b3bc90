#!/bin/bash
# one single-GPU round: all GPU tests, the default bench, A/B of the transfer-kernel variants, Gauss-Seidel family benches.
# bash tools/gpu_round.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -25 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f e2e %.3e"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()}, d.get("trisolve_finest"), d["config"].get("defect"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
timeout 900 python bench.py > $out/${tag}_bench_p1.json 2> $out/${tag}_bench_p1.err; summ $out/${tag}_bench_p1.json p1
for v in "k1 UGGPU_TR_K_INTERP=1 UGGPU_TR_K_RESTRICT=1" "k2 UGGPU_TR_K_INTERP=2 UGGPU_TR_K_RESTRICT=2" "k4 UGGPU_TR_K_INTERP=4 UGGPU_TR_K_RESTRICT=4" "nofixed UGGPU_NO_FIXED_WIDTH=1"; do
  set -- $v; name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu --steps 6 --e2e-steps 1 > $out/${tag}_ab_$name.json 2>&1; summ $out/${tag}_ab_$name.json $name
done
for sm in gs sgs; do
  timeout 600 python bench.py --no-cpu --steps 4 --e2e-steps 1 --top 6 --smoother $sm > $out/${tag}_bench_$sm.json 2>&1; summ $out/${tag}_bench_$sm.json $sm-257
done
timeout 400 python bench.py --no-cpu --steps 6 --e2e-steps 1 --top 6 > $out/${tag}_bench_jac257.json 2>&1; summ $out/${tag}_bench_jac257.json jac-257
