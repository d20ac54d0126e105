#!/bin/bash
# quick per-kernel counters (not a timing run): tools/ncu_quick.sh <kernel regex> <out csv> [extra bench args]
K=$1; OUT=$2; shift 2
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:$K -s 8 -c 3 --csv --log-file $OUT python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-extra "$@" > /dev/null 2>&1
