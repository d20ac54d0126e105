#!/usr/bin/env python
"""Debug: per-CTA phase times of train_tail_kernel (kodb_debug_set_train_trace hook)."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from object_detection_cib_b200 import _host as H
from object_detection_cib_b200 import _lib as L
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.iou import IoUCalculator, IoUType
from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams, launch_train_tail
from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

B, S, C = 64, 640, 10
dev = torch.device("cuda:0")
NSETS = 3  # rotate inputs and gradient buffers like bench.py does (3 x ~200 MB > L2)
targets = syn.make_targets(B, S, C, seed=2023)
mod = Yolov5Loss(Yolov5LabelAssigner(default_anchor_info()), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou),
                 syn.readme_class_weights())
plans, gradsets = [], []
for k in range(NSETS):
    heads = syn.make_head(B, S, C, seed=2023 + k, device=dev)
    net = [DetectionHeadResult(*lv) for lv in heads]
    plan = mod.plan(FeatureShape(S, S), net, targets)
    plans.append(plan)
    gradsets.append([H.alloc_planar_like(t) for t in plan.heads.tensors])
losses = torch.zeros(3, device=dev)
nass = torch.zeros(3, dtype=torch.int32, device=dev)
lib = L.lib()
lib.kodb_debug_set_train_trace.argtypes = [ctypes.c_void_p]
NG = 148 * 8 + B * 3
trace = torch.zeros(NG * 8, dtype=torch.long, device=dev)
for it in range(7):
    if it == 6:
        lib.kodb_debug_set_train_trace(trace.data_ptr())
    launch_train_tail(plans[it % NSETS], (float(B),) * 3, None, 0, gradsets[it % NSETS], losses, nass)
torch.cuda.synchronize()
lib.kodb_debug_set_train_trace(None)
t = trace.cpu().numpy().reshape(NG, 8).astype(np.float64)
used = t[:, 0] > 0
t0 = t[used, 0].min()
nd = int(used.sum()) - B * 3  # dense CTAs come first in block order
d = t[:nd]
sp = t[nd:nd + B * 3]
us = lambda a: (a - t0) / 1e3
print(f"dense CTAs {nd}: start {us(d[:, 0]).min():.1f}..{us(d[:, 0]).max():.1f}  end {us(d[:, 1]).min():.1f}..{us(d[:, 1]).max():.1f} us")
print(f"sparse CTAs {B * 3}: start {us(sp[:, 0]).min():.1f}..{us(sp[:, 0]).max():.1f}")
for nm, k in (("GT+scan", 1), ("gathers+iou", 4), ("dup scan", 5), ("cls", 2), ("published", 3), ("dense seen", 6), ("patched", 7)):
    v = us(sp[:, k])
    v = v[v > -1e6]
    print(f"  sparse {nm:12s} min {v.min():6.1f} mean {v.mean():6.1f} p90 {np.percentile(v, 90):6.1f} max {v.max():6.1f}")
fin = sp[sp[:, 2] > sp[:, 3]]
if len(fin):
    print(f"finalising sparse CTA: published {us(fin[:, 3]).max():.1f}, losses out at {us(fin[:, 2]).max():.1f} us")
print("tail after the last dense CTA ended: %.1f us" % (us(sp[:, 7]).max() - us(d[:, 1]).max()))
print("losses", losses.tolist(), "N_l", nass.tolist())
