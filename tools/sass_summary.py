#!/usr/bin/env python
"""SASS evidence per kernel of libkodb.so (sm_100a): instruction count and the mnemonics that matter
(tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, mbarrier -> SYNCS, cp.async -> LDGSTS, 128-bit global access).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "object_detection_cib_b200", "csrc")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UTMACCTL", "SYNCS", "LDGSTS", "LDG.128", "STG.128",
        "MUFU", "REDG", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "MATCH", "BAR", "DFMA", "DADD"]
ins = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)")
print("# cuobjdump -sass of the objects linked into libkodb.so (nvcc -gencode arch=compute_100a,code=sm_100a); per kernel: SASS")
print("# instructions and counts of selected mnemonics (prefix match; LDG.128 / STG.128: any 128-bit global load / store)")
for obj in ["head.o", "train_tail.o", "eval.o", "ap.o", "sgd.o", "iou.o", "assign.o", "decode.o"]:
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    fn, cnt, n = None, collections.Counter(), 0

    def flush():
        if fn:
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            sel = " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k])
            print(f"{obj:13s} {name[:64]:64s} {n:6d} instr | {sel}")
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            fn, cnt, n = m.group(1), collections.Counter(), 0
            continue
        m = ins.match(line)
        if m:
            n += 1
            op = m.group(1)
            for k in KEYS:
                if k in ("LDG.128", "STG.128"):
                    if op.startswith(k[:3] + ".") and "128" in op:
                        cnt[k] += 1
                elif op.startswith(k):
                    cnt[k] += 1
    flush()
