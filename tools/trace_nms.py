#!/usr/bin/env python
"""Debug: per-image phase times of select_nms_kernel (uses the kodb_debug_set_nms_trace hook)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from object_detection_cib_b200 import _lib as L
from object_detection_cib_b200 import nms as N
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

B, S, C = 64, 640, 10
dev = torch.device("cuda:0")
heads = syn.make_head(B, S, C, seed=2023, device=dev)
targets = syn.make_targets(B, S, C, seed=2023)
syn.inject_ground_truth(heads, targets, S, 0.3)
if len(sys.argv) > 1:  # detector-like clutter: trace_nms.py <radius>
    syn.light_up_neighbourhoods(heads, targets, radius=int(sys.argv[1]))
net = [DetectionHeadResult(*lv) for lv in heads]
gt_count = torch.zeros(C, dtype=torch.long, device=dev)
lib = L.lib()
trace = torch.zeros(B * 128, dtype=torch.long, device=dev)
lib.kodb_debug_set_nms_trace.argtypes = [ctypes.c_void_p]
for it in range(3):
    if it == 2:
        lib.kodb_debug_set_nms_trace(trace.data_ptr())
    res, rec = N.eval_tail(FeatureShape(S, S), net, default_anchor_info(), 0.001, 0.6, targets=targets, gt_count=gt_count)
torch.cuda.synchronize()
lib.kodb_debug_set_nms_trace(None)
# wall time of the two kernels of one call (CUDA events), same inputs, for comparison with the in-kernel stamps
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
times = []
for it in range(10):
    ev[0].record()
    N.eval_tail(FeatureShape(S, S), net, default_anchor_info(), 0.001, 0.6, targets=targets, gt_count=gt_count)
    ev[1].record()
    torch.cuda.synchronize()
    times.append(ev[0].elapsed_time(ev[1]) * 1e3)
print("eval_tail call (python launch + both kernels) us: min", min(times), "median", sorted(times)[5])
t = trace.cpu().numpy().reshape(B, 128)
names = {0: "start", 1: "tileprefix", 2: "select", 3: "collect", 4: "sort", 5: "decode", 6: "nms_compact", 7: "output", 8: "match", 10: "nms_bucket", 11: "nms_scatter", 12: "nms_tiles", 13: "nms_resolve"}
print("ncand", res.num_candidates[:8].tolist(), "kept", res.counts[:8].tolist(), "completions", lib.kodb_debug_take_completions())
print("select kernel: first CTA start -> last CTA end (globaltimer) us", (t[:, 127].max() - t[:, 1].min()) / 1e3, "per-CTA span mean/max us", ((t[:, 127] - t[:, 1]) / 1e3).mean(), ((t[:, 127] - t[:, 1]) / 1e3).max(), "start spread us", (t[:, 1].max() - t[:, 1].min()) / 1e3)
tot = {}
for b in range(B):
    n = int(t[b, 0])
    prev = None
    for k in range(n):
        pid, clk = int(t[b, 2 + 2 * k]), int(t[b, 3 + 2 * k])
        if prev is not None:
            tot.setdefault(names[pid], []).append((clk - prev) / 1965.0)
        prev = clk
    if b < 3:
        print(b, [(names[int(t[b, 2 + 2 * k])]) for k in range(n)], "total us", (int(t[b, 3 + 2 * (n - 1)]) - int(t[b, 3])) / 1965.0)
for k, v in tot.items():
    print(f"{k:10s} n/img={len(v) / B:4.1f} mean={np.mean(v):7.1f}us max={np.max(v):7.1f}us sum/img={np.sum(v) / B:7.1f}us")
