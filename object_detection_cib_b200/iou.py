"""`IoUType` / `IoUCalculator` configuration carriers (kod/core/bbox/iou.py:9-14, :249-268).

In the reference `IoUCalculator.__call__` runs the aligned IoU family as a chain of ATen elementwise
ops; here the calculator only carries `(iou_type, eps)` into the fused CUDA train-tail kernel
(csrc/train_tail.cu::iou_metric), which evaluates the same expression and its analytic gradient.
"""
from __future__ import annotations

import enum


@enum.unique
class IoUType(str, enum.Enum):
    ioU = "iou"
    giou = "giou"
    diou = "diou"
    ciou = "ciou"


class IoUCalculator(object):
    def __init__(self, iou_type: IoUType = IoUType.ciou, eps: float = 1e-7):
        self.iou_type = IoUType(iou_type)
        self.eps = float(eps)

    def __call__(self, boxes1, boxes2):
        raise RuntimeError(
            "IoUCalculator is evaluated inside the fused CUDA train tail (Yolov5Loss.forward); "
            "it is not a standalone op in object_detection_cib_b200.")
