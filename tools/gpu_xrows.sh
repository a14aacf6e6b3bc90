#!/bin/bash
# bash tools/gpu_xrows.sh <tag> <k>...: default bench with the exception-row kernels capped at k CTAs per SM (0 = one thread per row)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for k in "$@"; do
  UGGPU_XROWS_CTAS=$k python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 0 > $out/${tag}_x$k.json 2> $out/${tag}_x$k.err
  python - <<PY
import json
d = json.loads(open("$out/${tag}_x$k.json").read().strip().splitlines()[-1])
print("k=$k", round(d["ms_per_step"], 3), "ms/cycle; pair", round(d["roofline"]["avg_ms"], 4), "ms; spmv", round(d["spmv"]["avg_ms"], 4), "defect", d["config"]["defect"][1])
PY
done
