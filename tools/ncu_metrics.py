#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report.  usage: ncu_metrics.py report.ncu-rep [kernel_regex]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
kre = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
ki = h.index("Kernel Name")
for r in rows[2:]:
    if kre and not kre.search(r[ki]):
        continue
    print("Kernel", r[ki][:90])
    for k in want:
        if k in h:
            print(f"  {k:70s} {r[h.index(k)]:>16s} {rows[1][h.index(k)]}")
