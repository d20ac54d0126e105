"""Oracle (test infrastructure): torch-CPU restatement of the eval-side head decode.

Restates `DefaultYolov5Experiment.get_detections` (kod/lightning/experiments/yv5_baseline/exp.py:70-102)
= per level `Yolov5BoxPrediction.forward` (layers.py:55-63), `Yolov5ObjectnessPrediction` (:66-76),
`Yolov5ClassPrediction` (:79-89), then `Yolov5PredictionAssembler` (:127-155).

Row order inside an image: levels ll | ml | hl; inside a level r = (a*H + y)*W + x.
Columns: x1, y1, x2, y2 (pixels), sigmoid(obj), sigmoid(cls_0..C-1).
Unlike the reference this does not overwrite the box head tensor in place (callers never read it
afterwards, exp.py:146-154).  Pinned by tests/golden/decode_*.npz.
"""
from __future__ import annotations

from typing import Sequence

import torch


def decode_level(box: torch.Tensor, obj: torch.Tensor, cls: torch.Tensor, stride: int,
                 anchors_px: Sequence[Sequence[float]]):
    B, A, H, W, _ = box.shape
    ys, xs = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    # the reference builds grid and anchors on the CPU and moves them to the head's device (layers.py:57-60); run
    # with CUDA head tensors this executes torch's CUDA elementwise kernels, i.e. what the reference does on a GPU
    grid = torch.stack((xs, ys), 2).view(1, 1, H, W, 2).to(box.device)
    # torch.tensor(list of int pairs) is int64 in the reference; the product promotes to fp32
    anc = torch.tensor([[a[0], a[1]] for a in anchors_px]).reshape(1, A, 1, 1, 2).to(box.device)
    xy = box[..., 0:2].sigmoid() * 2 + grid - 0.5
    xy = xy * stride
    wh = ((box[..., 2:4].sigmoid() * 2) ** 2) * anc
    c = torch.cat((xy, wh), -1).reshape(B, -1, 4)
    cx, cy, w, h = c.unbind(-1)
    xyxy = torch.stack((cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h), -1)
    return xyxy, obj.sigmoid().reshape(B, -1, 1), cls.sigmoid().reshape(B, -1, cls.shape[-1])


def get_detections(heads, levels) -> torch.Tensor:
    """heads = [(box,obj,cls)] for ll, ml, hl; levels = [(stride, anchors_px)] -> (B, AC, 5+C)."""
    parts = [decode_level(b, o, c, s, a) for (b, o, c), (s, a) in zip(heads, levels)]
    boxes = torch.cat([p[0] for p in parts], 1)
    objs = torch.cat([p[1] for p in parts], 1)
    clss = torch.cat([p[2] for p in parts], 1)
    return torch.cat((boxes, objs, clss), -1)
