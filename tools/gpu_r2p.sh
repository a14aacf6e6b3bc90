#!/bin/bash
# ncu captures: bash tools/gpu_r2p.sh <tag> <kernel-regex> <skip> <count>
tag=$1; rx=$2; skip=$3; cnt=$4; out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $out/${tag} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 0 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}.ncu-rep --page details > $out/${tag}_details.txt 2>&1
ncu -i $out/${tag}.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread > $out/${tag}_raw.csv 2>&1
cat $out/${tag}_raw.csv | cut -c1-400 | tail -8
