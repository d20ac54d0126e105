#!/bin/bash
# one --set full capture (with source) of one kernel of the bench step: tools/ncu_kernel.sh <kernel regex> <out.ncu-rep>
K=$1; OUT=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 1 -o $OUT -f python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-extra "$@" > /dev/null 2>&1
