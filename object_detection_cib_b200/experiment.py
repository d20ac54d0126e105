"""Train/val step wiring of `DefaultYolov5Experiment` (kod/lightning/experiments/yv5_baseline/exp.py:36-185)
for the CUDA detection tail.

`Yolov5TailMixin` overrides exactly the three methods on the hot path – `get_detections` (:70-102),
`training_step` (:104-138), `validation_step` (:140-154) – and leaves the network, optimizer and
warm-up hooks to the reference class:

    class Experiment(Yolov5TailMixin, DefaultYolov5Experiment): ...

It expects the attributes the reference experiment already has: `self.net`, `self.loss` (our
`Yolov5Loss`), `self.anchor_info`, `self.val_nms_conf_threshold`, `self.val_nms_iou_threshold`,
`self.log`.
"""
from __future__ import annotations

import torch

from . import nms as N
from .types import FeatureShape


class Yolov5TailMixin(object):
    #: when set (a DeviceCocoMapEvaluator), validation_step also runs the COCO matching in the fused kernel
    tail_evaluator = None
    #: "list" reproduces the reference's return value (one host sync per batch); "padded" avoids the sync
    tail_detections_format = "list"

    def get_detections(self, image_feature_shape: FeatureShape, net_result) -> torch.Tensor:
        return N.get_detections(image_feature_shape, net_result, self.anchor_info)

    def training_step(self, batch, batch_idx: int):
        images, targets, _ = batch
        net_result = self.net(images)
        image_feature_shape = FeatureShape(width=images.shape[3], height=images.shape[2])
        loss_result = self.loss(image_feature_shape, net_result, targets)
        batch_size = images.shape[0]
        total_loss = batch_size * (loss_result.localization + loss_result.classification + loss_result.objectness)
        self.log("obj", loss_result.objectness, prog_bar=True, batch_size=batch_size)
        self.log("cls", loss_result.classification, prog_bar=True, batch_size=batch_size)
        self.log("box", loss_result.localization, prog_bar=True, batch_size=batch_size)
        return total_loss

    def validation_step(self, batch, batch_idx: int):
        images, targets, _ = batch
        net_result = self.net(images)  # raw head maps; decode + NMS are fused below
        image_feature_shape = FeatureShape(width=images.shape[3], height=images.shape[2])
        ev = self.tail_evaluator
        out = N.eval_tail(image_feature_shape, net_result, self.anchor_info,
                          conf_thres=self.val_nms_conf_threshold, nms_thres=self.val_nms_iou_threshold,
                          targets=targets if ev is not None else None,
                          gt_count=ev.gt_count if ev is not None else None)
        res, records = out if ev is not None else (out, None)
        detections = N.unpad(res) if self.tail_detections_format == "list" else res
        return {"targets": targets, "detections": detections, "records": records}
