#!/bin/bash
# bash tools/gpu_asm.sh <tag>: the assembly and Galerkin setup operations at size (bench.py's own functions, one GPU)
tag=$1; out=gpurun_out; mkdir -p $out
python - > $out/${tag}_asm.json 2> $out/${tag}_asm.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
for top in (5, 6):
    print(json.dumps(bench.assemble_bench(0, (4, 4, 4), top)))
PY
cat $out/${tag}_asm.json; tail -3 $out/${tag}_asm.err
