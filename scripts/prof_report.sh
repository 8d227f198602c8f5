#!/bin/bash
# scripts/prof_report.sh <ncu-rep> <kernel symbol fragment> [top_n]: per-line report of a capture (run here, no GPU)
rep=$1; sym=$2; top=${3:-40}
tmp=/tmp/frxprof; mkdir -p $tmp/elf; rm -f $tmp/elf/*
(cd $tmp/elf && cuobjdump -xelf all /root/repo/frenetix_motion_planner_b200/libfrx_b200.so > /dev/null && nvdisasm --print-line-info-inline frx_kernels.sm_100a.cubin > k.dis 2>/dev/null)
ncu -i $rep --page source --csv > $tmp/src.csv 2>/dev/null
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','launch__shared_mem_per_block_dynamic']
for k,x in zip(h,v):
    if k in want: print(k, x, rows[1][h.index(k)])
"
python3 /root/repo/profiles/ncu_by_line.py $tmp/src.csv $tmp/elf/k.dis $sym $top
