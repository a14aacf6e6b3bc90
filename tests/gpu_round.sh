#!/bin/bash
# One GPU-box visit: parity tests, bench A/B, ncu captures.  Usage: gpurun -- bash tests/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest.log
python bench.py --no-cpu > $out/${tag}_bench_p1.log 2>&1
UGGPU_NO_COL_COMPRESSION=1 python bench.py --no-cpu > $out/${tag}_bench_p1_nocomp.log 2>&1
python bench.py --kind q1 --no-cpu > $out/${tag}_bench_q1.log 2>&1
python bench.py --kind elasticity --top 6 --no-cpu > $out/${tag}_bench_el.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_smooth_k -s 32 -c 3 -o $out/${tag}_smooth -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_smooth.ncu-rep --page details > $out/${tag}_smooth_details.txt 2>&1
for f in $out/${tag}_pytest.log $out/${tag}_bench_*.log; do echo "== $f"; tail -c 600 $f; echo; done
