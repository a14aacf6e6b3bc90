#!/bin/bash
# bash tools/gpu_r2x.sh <tag>: what the driver runs at round end -- GPU tests, smoke, the default bench line, the reference arm
tag=$1; out=gpurun_out; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $out/${tag}_pytest.log 2>&1; tail -4 $out/${tag}_pytest.log
( time python __graft_entry__.py smoke ) > $out/${tag}_smoke.log 2>&1; tail -4 $out/${tag}_smoke.log
( time python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $out/${tag}_ref.json 2> $out/${tag}_ref.err; tail -c 400 $out/${tag}_ref.json; tail -4 $out/${tag}_ref.err
