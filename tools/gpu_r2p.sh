#!/bin/bash
# ncu captures: bash tools/gpu_r2p.sh <tag> <kernel-regex> <skip> <count>
tag=$1; rx=$2; skip=$3; cnt=$4; out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o $out/${tag} -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extras --e2e-steps 0 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}.ncu-rep --page details > $out/${tag}_details.txt 2>&1
ncu -i $out/${tag}.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__warps_eligible.avg.per_cycle_active > $out/${tag}_raw.csv 2>&1
rm -f $out/${tag}.ncu-rep
python - $out/${tag}_raw.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if hdr:
    h=rows[hdr[0]]; data=rows[hdr[0]+2:]
    ix={n:h.index(n) for n in h}
    for r in data:
        if len(r)<len(h): continue
        print(r[ix["Kernel Name"]][:60], r[ix["Grid Size"]], "t=%s"%r[ix["gpu__time_duration.sum"]], "rd=%s wr=%s"%(r[ix["dram__bytes_read.sum"]],r[ix["dram__bytes_write.sum"]]), "dram%%=%s L2hit=%s L1hit=%s occ=%s elig=%s regs=%s"%(r[ix.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",0)],r[ix["lts__t_sector_hit_rate.pct"]],r[ix["l1tex__t_sector_hit_rate.pct"]],r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]],r[ix["smsp__warps_eligible.avg.per_cycle_active"]],r[ix["launch__registers_per_thread"]]))
PY
