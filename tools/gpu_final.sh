#!/bin/bash
# bash tools/gpu_final.sh <tag>: tests, smoke, bench lines and ncu evidence of the round's final state (most important first)
tag=$1; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $out/${tag}_bench_p1.json 2> $out/${tag}_bench_p1.err; tail -c 200 $out/${tag}_bench_p1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_smooth_sten -s 32 -c 2 -o $out/${tag}_smooth -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_smooth.ncu-rep --page details > $out/${tag}_smooth_details.txt 2>&1
ncu -i $out/${tag}_smooth.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct > $out/${tag}_smooth_raw.csv 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_launches.log 2>&1
timeout 300 python bench.py --impl reference --steps 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 200 python bench.py --kind q1 --no-cpu > $out/${tag}_bench_q1.json 2>&1
timeout 200 python bench.py --kind elasticity --top 6 --no-cpu > $out/${tag}_bench_el.json 2>&1
timeout 200 python bench.py --smoother ilu --no-cpu --steps 3 --e2e-steps 1 > $out/${tag}_bench_ilu513.json 2>&1
for f in p1 q1 el ilu513; do python - $out/${tag}_bench_$f.json $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print(sys.argv[2], "%.3e unk/s %.2f ms/step; dom %.3f ms %.0f GB/s frac %.3f; cycle_frac %.3f e2e %.3e launches %s"%(d["value"],d["ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],r["cycle_frac"],d["e2e"]["value"],d["gpu_launches"]), {k:round(v["ms"]/d["steps"],2) for k,v in d["kernels"].items() if v["ms"]>0.05}, d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done
tail -c 400 $out/${tag}_bench_ref.json
