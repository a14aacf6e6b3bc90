#!/bin/bash
# bash tools/gpu_round2.sh <tag>: GPU tests, transfer-kernel A/B (prefetch placement, fixed width), ncu of the transfer kernels, GS benches
tag=$1; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -15 $out/${tag}_pytest.log
summ() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f e2e %.3e"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items()}, d.get("trisolve_finest"), d["config"].get("defect"))
except Exception as e:
    print(sys.argv[2], "failed", e); print(open(sys.argv[1]).read()[-1500:])
PY
}
for v in "early UGGPU_PF_MODE=63" "late UGGPU_PF_MODE=31" "nofixed UGGPU_NO_FIXED_WIDTH=1" "early2 UGGPU_PF_MODE=63" "nopfvec UGGPU_PF_MODE=59"; do
  set -- $v; name=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu --steps 6 --e2e-steps 1 > $out/${tag}_ab_$name.json 2>&1; summ $out/${tag}_ab_$name.json $name
done
for sm in gs sgs; do
  timeout 600 python bench.py --no-cpu --steps 4 --e2e-steps 1 --top 6 --smoother $sm > $out/${tag}_bench_$sm.json 2>&1; summ $out/${tag}_bench_$sm.json $sm-257
done
timeout 900 python bench.py --no-cpu --steps 3 --e2e-steps 1 --smoother gs > $out/${tag}_bench_gs513.json 2>&1; summ $out/${tag}_bench_gs513.json gs-513
ncu --set full --clock-control none --import-source on -k regex:"k_interpolate_k|k_restrict_k" -s 14 -c 2 -o $out/${tag}_transfer -f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_transfer.ncu-rep --page details > $out/${tag}_transfer_details.txt 2>&1
grep -E "k_interpolate_k|k_restrict_k|Duration|DRAM Throughput|L2 Hit|Achieved Occupancy|Registers Per|Eligible|dram__bytes" $out/${tag}_transfer_details.txt | head -40
