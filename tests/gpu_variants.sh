#!/bin/bash
# A/B of kernel build variants on one GPU box: bash tests/gpu_variants.sh <tag> <lib suffixes...>   ("" = default library)
tag=$1; shift
out=gpurun_out; mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
cat $out/${tag}_pytest.log
for v in "$@"; do
  lib=ug_b200/lib/libuggpu${v}.so
  for kind in p1 q1; do
    UGGPU_LIB=$PWD/$lib python bench.py --no-cpu --steps 5 --e2e-steps 1 --kind $kind > $out/${tag}_bench_${kind}${v}.log 2>&1
    python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_bench_${kind}${v}.log").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$kind$v", "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"]))
except Exception as e:
    print("$kind$v failed", e)
PY
  done
done
UGGPU_NO_COL_COMPRESSION=1 python bench.py --no-cpu --steps 5 --e2e-steps 1 > $out/${tag}_bench_p1_nocomp.log 2>&1; tail -c 300 $out/${tag}_bench_p1_nocomp.log
