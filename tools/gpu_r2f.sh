#!/bin/bash
# bash tools/gpu_r2f.sh <tag>: final state of the round -- what the driver runs (tests, smoke, bench, reference arm), the ncu launch list of the
# bench command and one --set full capture of the dominant kernel pair on the finest level
tag=$1; out=gpurun_out; mkdir -p $out
bash tools/gpu_r2x.sh $tag
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --e2e-steps 0 > $out/${tag}_launches.log 2>&1
bash tools/gpu_r2p.sh ${tag}_stx "k_smooth_stx|k_smooth_xrows" 60 4
