#!/usr/bin/env python
"""Turn the outputs of tools/profile_round.sh into the tracked files under profiles/:
   <tag>_summary.txt (key metrics of the --set full captures), <tag>_launches.csv (launch list), r1_traffic.json."""
import csv
import json
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2_final"
import os

rows = []
h = None
for rep in (f"gpurun_out/{tag}_full.ncu-rep", f"gpurun_out/{tag}_head.ncu-rep"):
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    if h is None:
        h, rows = rr[0], rr[:2]
    hh = rr[0]
    for r in rr[2:]:  # re-order the columns of a second report to the first one's header
        rows.append([r[hh.index(k)] if k in hh else "" for k in h])
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"]
ki = h.index("Kernel Name")
traffic = {}
with open(f"profiles/{tag}_summary.txt", "w") as f:
    f.write(f"# from {tag}_full.ncu-rep: ncu --set full --clock-control none --import-source on, one launch per kernel,\n"
            f"# command: python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline (B=64, 640x640, 10 classes)\n")
    for r in rows[2:]:
        name = r[ki]
        f.write(f"\nKernel Name  {name}\n")
        vals = {}
        for k in want:
            if k in h:
                i = h.index(k)
                f.write(f"{k:62s} {r[i]:>18s} {rows[1][i]}\n")
                vals[k] = (r[i], rows[1][i])

        def mb(key):
            v, u = vals[key]
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        short = ("train_tail_kernel" if "train_tail" in name else "cand_floor_kernel" if "cand_floor" in name else
                 "head_forward_kernel" if "head_forward" in name else "select_nms_kernel")
        traffic[short] = int(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
json.dump({"source": f"profiles/{tag}_summary.txt (ncu --set full --clock-control none), dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "images_per_launch": 64, "kernels": traffic}, open("profiles/r2_traffic.json", "w"), indent=2)
shutil.copy(f"gpurun_out/{tag}_launches.csv", f"profiles/{tag}_launches.csv")
print(open(f"profiles/{tag}_summary.txt").read())
print(subprocess.run([sys.executable, "tools/launch_summary.py", f"profiles/{tag}_launches.csv"], capture_output=True, text=True).stdout)
