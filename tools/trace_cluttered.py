#!/usr/bin/env python
"""How the eval tail behaves on detector-like output: every ground-truth box lights up a neighbourhood of cells on all
three anchors (hundreds of overlapping high-score candidates per object), so NMS must scan thousands of candidates
for its 300 survivors.  Prints per-call time of the two eval kernels (CUDA events, warm) and the completion count."""
import sys

import torch

sys.path.insert(0, ".")
from object_detection_cib_b200 import _lib as L
from object_detection_cib_b200 import nms as N
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

B, S, C = 64, 640, 10
dev = torch.device("cuda:0")
radius = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n_gt = int(sys.argv[2]) if len(sys.argv) > 2 else 15
heads = syn.make_head(B, S, C, seed=2023, device=dev)
targets = syn.make_targets(B, S, C, seed=2023, n_min=n_gt, n_max=n_gt)
syn.light_up_neighbourhoods(heads, targets, radius=radius)
net = [DetectionHeadResult(*lv) for lv in heads]
lib = L.lib()
shape, info = FeatureShape(S, S), default_anchor_info()
det = N.get_detections(shape, net, info)
want = N.nms_padded(det, 0.001, 0.6)
ws_times = []
lib.kodb_debug_take_completions()
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    got = N.eval_tail(shape, net, info, 0.001, 0.6)
    e1.record()
    torch.cuda.synchronize()
    ws_times.append(e0.elapsed_time(e1) * 1e3)
print("radius", radius, "gt/img", n_gt, "candidates/img", got.num_candidates[:4].tolist(), "kept", got.counts[:4].tolist())
print("eval_tail call incl. host launch overhead (us):", [round(t) for t in ws_times], "completions", lib.kodb_debug_take_completions())
print("equal to kodb_nms(kodb_decode()):", bool(torch.equal(got.dets, want.dets) and torch.equal(got.keep_idx, want.keep_idx)))
