#!/usr/bin/env python
"""Device time of kodb_train_tail per launch: N back-to-back launches between one pair of CUDA events, queued behind a
spin kernel, inputs rotating over 3 sets (> L2), like bench.py's roofline leg.  Usage: time_train.py [B S C n_gt]"""
import sys

import torch

sys.path.insert(0, ".")
from object_detection_cib_b200 import _host as H
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.iou import IoUCalculator, IoUType
from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams, launch_train_tail
from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

B, S, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 640, 10)
n_gt = int(sys.argv[4]) if len(sys.argv) > 4 else -1
dev = torch.device("cuda:0")
NSETS, N = 3, 60
mod = Yolov5Loss(Yolov5LabelAssigner(default_anchor_info()), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou),
                 syn.readme_class_weights() if C == 10 else None)
plans, gradsets = [], []
for k in range(NSETS):
    heads = syn.make_head(B, S, C, seed=2023 + k, device=dev)
    targets = syn.make_targets(B, S, C, seed=2023 + k, n_min=max(n_gt, 0), n_max=n_gt if n_gt >= 0 else 10)
    plan = mod.plan(FeatureShape(S, S), [DetectionHeadResult(*lv) for lv in heads], targets)
    plans.append(plan._replace(workspace=H.workspace(f"train{k}", dev, plan.workspace.numel())))
    gradsets.append([H.alloc_planar_like(t) for t in plan.heads.tensors])
losses = torch.zeros(3, device=dev)
nass = torch.zeros(3, dtype=torch.int32, device=dev)
hint = (float(B),) * 3
for i in range(6):
    launch_train_tail(plans[i % NSETS], hint, None, 0, gradsets[i % NSETS], losses, nass)
torch.cuda.synchronize()
best = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(0.03 * 1.9e9))
    e0.record()
    for i in range(N):
        launch_train_tail(plans[i % NSETS], hint, None, 0, gradsets[i % NSETS], losses, nass)
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / N * 1e3)
ac = sum(3 * (S // s) ** 2 for s in (8, 16, 32))
nbytes = 4 * ac * (1 + 5 + C) * B
us = min(best)
print(f"train_tail B={B} S={S} C={C}: {us:.2f} us/launch ({[round(b, 2) for b in best]}), {nbytes / us / 1e3:.0f} GB/s, "
      f"frac of 6551 = {nbytes / us / 1e3 / 6551:.3f}, losses {losses.tolist()} N_l {nass.tolist()}")
