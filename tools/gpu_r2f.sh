#!/bin/bash
out=gpurun_out; mkdir -p $out
for v in fake1 fake2; do
  if [ $v = fake1 ]; then export UGGPU_DBG_FAKE_COMM=1; fi; if [ $v = fake2 ]; then export UGGPU_DBG_FAKE_COMM=2; fi
  timeout 300 python bench.py --no-extras --no-cpu --steps 10 --e2e-steps 0 > $out/r2f_$v.json 2> $out/r2f_$v.err
  python - $out/r2f_$v.json $v <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[2], "%.2f ms/step dom %.3f ms"%(d["ms_per_step"], r["avg_ms"]), {k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]>0})
PY
done
