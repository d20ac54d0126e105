#!/bin/bash
# one --set full capture of the head kernel: tools/ncu_head.sh <out.ncu-rep> [B S C]
OUT=$1; shift
ncu --set full --clock-control none --import-source on -k regex:head_forward -s 6 -c 1 -o $OUT -f python tools/bench_head.py "$@" > /dev/null 2>&1
