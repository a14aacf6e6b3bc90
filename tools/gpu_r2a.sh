#!/bin/bash
# round 2, first N=1 check: tests, smoke, bench with all extra lines.  bash tools/gpu_r2a.sh <tag>
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; tail -15 $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/${tag}_bench_p1.json 2> $out/${tag}_bench_p1.err; tail -c 600 $out/${tag}_bench_p1.err
python - $out/${tag}_bench_p1.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
    print("main %.3e unk/s %.2f ms/step (with events %.2f, kernel sum %.2f); dom %.3f ms %.0f GB/s frac %.3f; e2e %.3e e2e_solve %.3e launches %s"%(d["value"],d["ms_per_step"],d["config"]["ms_per_step_with_kernel_events"],d["config"]["kernel_sum_ms_per_step"],r["avg_ms"],r["achieved"],r["frac"],d["e2e"]["value"],d["e2e_solve"]["value"],d["gpu_launches"]))
    print({k:round(v["ms"]/d["steps"],3) for k,v in d["kernels"].items() if v["ms"]>0})
    print("spmv", d.get("spmv"))
    for k in ("shared_tables_generic_kernel","general_path","varying_coefficient","q1_poisson","elasticity_3x3","galerkin","krylov"):
        e=d.get(k,{})
        if "error" in e: print(k,"ERROR",e["error"]); continue
        dk=e.get("dominant_kernel",{})
        print(k, "%.3e unk/s %.2f ms; dom %s %.3f ms frac %.3f survey %.3f"%(e.get("value",0),e.get("ms_per_step",0),dk.get("kernel","")[:24],dk.get("avg_ms",0),dk.get("frac",0),dk.get("frac_survey_model",0)), e.get("kernels_ms_per_step"), e.get("spmv"), e.get("galerkin_ms_top_down"), e.get("galerkin_finest"), e.get("krylov"))
    print("cpu", d.get("cpu_baseline")); print("equal", d.get("equal_size_inside_ug"))
except Exception as ex:
    print("failed", ex); print(open(sys.argv[1]).read()[-2000:])
PY
