#!/bin/bash
# bash tools/gpu_kry.sh <tag>: Krylov parity (fused chains vs separate calls vs the reference's dumps) and the cg / bcgs bench key, fused and not
tag=$1; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "krylov" 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --krylov --e2e-steps 0 > $out/${tag}_kry_fused.json 2> $out/${tag}_kry_fused.err
UGGPU_NO_KRYLOV_FUSION=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --krylov --e2e-steps 0 > $out/${tag}_kry_plain.json 2> $out/${tag}_kry_plain.err
python - <<PY
import json
for f in ("fused", "plain"):
    d = json.loads(open("$out/${tag}_kry_%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], json.dumps(d.get("krylov")))
PY
