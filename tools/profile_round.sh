#!/bin/bash
# Round profile pass (run on the GPU box through gpurun): plain run first (must exit 0), then the ncu launch list of
# the same command, then one `--set full` capture of each of the three kernels of a step.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r2_final}
CMD="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'train_tail_kernel|cand_floor_kernel|select_nms_kernel' -s 12 -c 3 \
    -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
# the Detect head (extra_configs f1_detect_head runs after the tail workloads): one capture of its kernel
ncu --set full --clock-control none --import-source on -k regex:head_forward_kernel -s 2 -c 1 \
    -o gpurun_out/${TAG}_head python tools/bench_head.py > gpurun_out/${TAG}_ncu_head.log 2>&1
ls -la gpurun_out/${TAG}_*
