#!/bin/bash
tag=$1
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu > $out/${tag}_bench_p1.json 2> $out/${tag}_bench_p1.err; tail -c 300 $out/${tag}_bench_p1.err
python - $out/${tag}_bench_p1.json p1 <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f e2e %.3e (%.1f ms)"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()})
PY
ncu --set full --clock-control none --import-source on -k regex:k_interpolate_k --launch-skip 6 --launch-count 1 -o $out/${tag}_interp -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_restrict_k --launch-skip 0 --launch-count 1 -o $out/${tag}_restrict -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_jac_k --launch-skip 0 --launch-count 1 -o $out/${tag}_jac -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu3.log 2>&1
ls -la $out/${tag}_*.ncu-rep
