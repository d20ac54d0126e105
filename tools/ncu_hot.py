#!/usr/bin/env python
"""Top source lines by warp-stall samples from an ncu report (needs -lineinfo + --import-source on).

usage: ncu_hot.py report.ncu-rep kernel_regex [top_n] [launch_skip]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
fname = ""
acc = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ii = hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= si or not r[0]:
        continue
    try:
        s = float(r[si])
        inst = float(r[ii])
    except ValueError:
        continue
    stalls = sorted(((float(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    acc.append((s, inst, fname, r[0], r[1].strip(), stalls))
tot = sum(a[0] for a in acc) or 1
toti = sum(a[1] for a in acc) or 1
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for s, inst, f, ln, src, st in sorted(acc, reverse=True)[:top]:
    sts = " ".join(f"{h[6:]}={v:.0f}" for v, h in st if v)
    print(f"{s / tot * 100:5.1f}% inst {inst / toti * 100:5.1f}%  {f}:{ln:>4s}  {src[:95]}  [{sts}]")
