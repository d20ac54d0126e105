"""Boundary types, field-for-field the reference's NamedTuples, so either side's objects can be passed.

* FeatureShape            kod/core/types.py:6-8
* AnchorBoxInfo           kod/core/anchors/info.py:9-11
* DetectionTarget         kod/data/detection.py:24-26
* DetectionHeadResult     kod/nn/heads/types.py:8-11
* Yolov5NetworkResult     kod/nn/networks/yolov5.py:34-37
* LayerwiseAnchorInfo / LossResult   kod/lightning/experiments/yv5_baseline/type_defs.py:16-31
* Assignment*             kod/core/label_assignment/yv5.py:18-42
* MAPEvalCallbackInput    kod/lightning/callbacks/map_eval.py:13-15
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import torch


class FeatureShape(NamedTuple):
    width: int
    height: int


class AnchorBoxInfo(NamedTuple):
    stride: int
    boxes_wh: Sequence[FeatureShape]


class DetectionTarget(NamedTuple):
    boxes: torch.Tensor
    labels: torch.Tensor


class DetectionHeadResult(NamedTuple):
    box: torch.Tensor
    obj: torch.Tensor
    cls: torch.Tensor


class Yolov5NetworkResult(NamedTuple):
    ll: DetectionHeadResult
    ml: DetectionHeadResult
    hl: DetectionHeadResult


class LayerwiseAnchorInfo(NamedTuple):
    ll: AnchorBoxInfo
    ml: AnchorBoxInfo
    hl: AnchorBoxInfo


AssignmentAnchorInfo = LayerwiseAnchorInfo  # same three fields (yv5.py:18-21)


class LossResult(NamedTuple):
    localization: torch.Tensor
    objectness: torch.Tensor
    classification: torch.Tensor


class AssignmentTargetIndices(NamedTuple):
    samples: torch.Tensor
    anchors: torch.Tensor
    grid_y: torch.Tensor
    grid_x: torch.Tensor


class AssignmentTargetInfo(NamedTuple):
    indices: AssignmentTargetIndices
    labels: torch.Tensor
    gt_boxes: torch.Tensor
    anchors: torch.Tensor
    target_feature_shape: FeatureShape


class AssignmentResult(NamedTuple):
    ll: AssignmentTargetInfo
    ml: AssignmentTargetInfo
    hl: AssignmentTargetInfo


def default_anchor_info() -> LayerwiseAnchorInfo:
    """configs/anchor_boxes/voc_s{8,16,32}.yaml (== kod/test_utils/anchor_boxes.py:6-31)."""
    from .synthetic import ANCHORS_PX, STRIDES

    return LayerwiseAnchorInfo(*[
        AnchorBoxInfo(stride=s, boxes_wh=[FeatureShape(width=w, height=h) for (w, h) in a])
        for s, a in zip(STRIDES, ANCHORS_PX)
    ])
