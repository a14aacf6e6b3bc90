#!/bin/bash
# bash tools/gpu_zero.sh <tag>: A/B of the zero-coefficient skipping (stencil kernels + packed exception rows), P1 / Q1 / elasticity, and the tests of the row forms
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_synth.py tests/test_gpu_parity.py -q -x -m gpu 2>&1 | tail -3
for kind in p1 q1 elasticity; do
  top=7; [ $kind = elasticity ] && top=6
  for keep in 0 1; do
    if [ $keep = 1 ]; then export UGGPU_KEEP_ZERO_ENTRIES=1; else unset UGGPU_KEEP_ZERO_ENTRIES; fi
    python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --e2e-steps 0 --kind $kind --top $top > $out/${tag}_${kind}_keep$keep.json 2> $out/${tag}_${kind}_keep$keep.err
    python - <<PY
import json
d = json.loads(open("$out/${tag}_${kind}_keep$keep.json").read().strip().splitlines()[-1])
sp = d.get("spmv") or {}
print("$kind keep_zeros=$keep", round(d["ms_per_step"], 3), "ms/cycle; pair", round(d["roofline"]["avg_ms"], 4), "ms, frac", round(d["roofline"]["frac"], 3), "alg GB", round(d["roofline"]["alg_bytes_per_launch"] / 1e9, 3), "; spmv", sp.get("avg_ms"), "defect", d["config"]["defect"])
PY
  done
done
