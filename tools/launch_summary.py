#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v if u.startswith("ns") or u == "nsecond" else v * 1e3 if u.startswith("us") else v * 1e6 if u.startswith("ms") else v
        agg[row["Kernel Name"][:60]].append(v)
tot = sum(sum(v) for v in agg.values())
only = sys.argv[2] if len(sys.argv) > 2 else ""
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if only and only not in k:
        continue
    print(f"{k:62s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:8.1f}us min={min(v) / 1e3:8.1f}us share={sum(v) / tot * 100:5.1f}%")
