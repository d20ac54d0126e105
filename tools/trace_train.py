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
nd = int(used.sum()) - B * 3
d = t[:nd]
sp = t[nd:nd + B * 3]
print("dense CTAs", nd, "start min/max us", (d[:, 0].min() - t0) / 1e3, (d[:, 0].max() - t0) / 1e3, "end min/max", (d[:, 1].min() - t0) / 1e3,
      (d[:, 1].max() - t0) / 1e3)
names = ["start", "N_l", "rows", "waited", "patched", "partials", "final"]
for k in range(6):
    v = (sp[:, k] - t0) / 1e3
    print(f"sparse {names[k]:9s} min={v.min():7.1f} mean={v.mean():7.1f} max={v.max():7.1f} us")
fin = sp[:, 6].max()
print("finaliser end us", (fin - t0) / 1e3, "losses", losses.tolist())
