#!/usr/bin/env python
"""Per-source-line executed warp instructions (per warp) from an ncu report.
usage: ncu_lines.py report.ncu-rep kernel_regex n_warps [top_n]"""
import csv, io, subprocess, sys
rep, kern, nw = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{kern}", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
acc = []
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        ti = hdr.index("Thread Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or not r[0] or len(r) <= ii:
        continue
    try:
        acc.append((float(r[ii]), r[0], r[1].strip()[:105], float(r[ti]), float(r[si])))
    except ValueError:
        pass
tot = sum(a[0] for a in acc)
ts = sum(a[4] for a in acc) or 1
print("total", tot, "per warp", tot / nw)
for inst, ln, src, ti, sa in sorted(acc, reverse=True)[:top]:
    print(f"{inst / nw:7.1f} ({inst / tot * 100:4.1f}%) act {ti / max(inst, 1):4.1f} smp {sa / ts * 100:4.1f}% L{ln}: {src}")
