"""Drop-in `Yolov5LabelAssigner` (kod/core/label_assignment/yv5.py:45-319) on the CUDA kernel
`kodb_assign` (csrc/assign.cu): same constructor, same `__call__(input_image_shape, targets) ->
AssignmentResult`, rows in the reference's order, int64 indices bit-identical."""
from __future__ import annotations

from typing import Sequence

import torch

from . import _host as H
from . import _lib as L
from .types import (AssignmentAnchorInfo, AssignmentResult, AssignmentTargetIndices, AssignmentTargetInfo,
                    FeatureShape)


class Yolov5LabelAssigner(object):
    def __init__(self, anchor_info: AssignmentAnchorInfo, threshold: float = 4.0, device=None):
        self.anchor_info = anchor_info
        self.threshold = threshold
        self.off_bias = 0.5
        self.device = device

    def assign_padded(self, input_image_shape: FeatureShape, targets: Sequence, device=None):
        """No-sync form: per level the seven output tensors at capacity 5*A*T plus a device count."""
        device = torch.device(device or self.device or _first_device(targets) or "cuda")
        lib = L.lib()
        packed = H.pack_targets(targets, device)
        levels = H.level_cfgs(self.anchor_info, input_image_shape.width, input_image_shape.height)
        amax = max(lv.num_anchors for lv in levels)
        cap = max(5 * amax * packed.num_targets, 1)
        ints = torch.empty((L.NUM_LEVELS, 5, cap), dtype=torch.long, device=device)
        gtb = torch.empty((L.NUM_LEVELS, cap, 4), dtype=torch.float32, device=device)
        anc = torch.empty((L.NUM_LEVELS, cap, 2), dtype=torch.float32, device=device)
        counts = torch.zeros(L.NUM_LEVELS, dtype=torch.int32, device=device)
        outs = (L.AssignOut * L.NUM_LEVELS)()
        for i in range(L.NUM_LEVELS):
            o = outs[i]
            o.samples, o.anchors, o.grid_y, o.grid_x, o.labels = (ints[i, j].data_ptr() for j in range(5))
            o.gt_boxes, o.anchor_wh = gtb[i].data_ptr(), anc[i].data_ptr()
        L.check(lib.kodb_assign(levels, packed.struct, float(self.threshold), outs, cap, counts.data_ptr(),
                                H.stream_ptr(device)), "kodb_assign")
        return ints, gtb, anc, counts, levels, packed

    def __call__(self, input_image_shape: FeatureShape, targets: Sequence) -> AssignmentResult:
        ints, gtb, anc, counts, levels, _ = self.assign_padded(input_image_shape, targets)
        n = counts.tolist()  # the reference API returns exact-length tensors: one sync
        res = []
        for i in range(L.NUM_LEVELS):
            k = n[i]
            res.append(AssignmentTargetInfo(
                indices=AssignmentTargetIndices(samples=ints[i, 0, :k], anchors=ints[i, 1, :k],
                                                grid_y=ints[i, 2, :k], grid_x=ints[i, 3, :k]),
                labels=ints[i, 4, :k], gt_boxes=gtb[i, :k], anchors=anc[i, :k],
                target_feature_shape=FeatureShape(width=levels[i].W, height=levels[i].H)))
        return AssignmentResult(*res)


def _first_device(targets):
    for t in targets:
        if t[0].is_cuda:
            return t[0].device
    return None
