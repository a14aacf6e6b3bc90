#!/bin/bash
# bash tools/gpu_ytile.sh <tag> <k>...: P1 and Q1 bench with k passes per warp in the scalar stencil-row kernel (UGGPU_STX_YTILE)
tag=$1; shift; out=gpurun_out; mkdir -p $out
for kind in p1 q1; do
  for k in "$@"; do
    UGGPU_STX_YTILE=$k python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 0 --kind $kind > $out/${tag}_${kind}_y$k.json 2> $out/${tag}_${kind}_y$k.err
    python - <<PY
import json
d = json.loads(open("$out/${tag}_${kind}_y$k.json").read().strip().splitlines()[-1])
print("$kind ytile=$k", round(d["ms_per_step"], 3), "ms/cycle; pair", round(d["roofline"]["avg_ms"], 4), "ms frac", round(d["roofline"]["frac"], 3), "defect", d["config"]["defect"][1])
PY
  done
done
