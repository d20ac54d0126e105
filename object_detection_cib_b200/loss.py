"""Drop-in `Yolov5Loss` / `Yolov5LossParams` (kod/lightning/experiments/yv5_baseline/loss.py:25-248)
on the fused CUDA train tail `kodb_train_tail` (csrc/train_tail.cu).

Same constructor `(assigner, hparams, iou_calculator, weights)`, same
`forward(image_feature_shape, net_result, targets) -> LossResult` of three 0-d tensors that carry
autograd to the nine head tensors.  The forward launch already writes the gradients for the hinted
upstream scale (`training_step` multiplies the sum by the batch size, exp.py:124-130); backward
re-launches the kernel with the actual upstream gradients on the device, and the kernel exits
immediately when they equal the hint – no host sync either way.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence

import torch
import torch.nn as nn

from . import _host as H
from . import _lib as L
from .iou import IoUCalculator
from .label_assignment import Yolov5LabelAssigner
from .types import FeatureShape, LossResult


class Yolov5LossParams(NamedTuple):
    lambda_classification: float
    lambda_localization: float
    lambda_objectness: float

    lambda_ll_objectness: float
    lambda_ml_objectness: float
    lambda_hl_objectness: float

    @staticmethod
    def get_default() -> "Yolov5LossParams":
        return Yolov5LossParams(0.5, 0.05, 1.0, 4.0, 1.0, 0.4)


class TrainTailPlan(NamedTuple):
    """Everything one launch of kodb_train_tail needs (built once per step on the host)."""

    heads: H.Heads
    levels: object
    packed: H.PackedTargets
    cfg: L.LossCfg
    pos_weight: Optional[torch.Tensor]
    workspace: torch.Tensor


def launch_train_tail(plan: TrainTailPlan, upstream_h, upstream_dev, skip_if_equal, grads, losses, n_assigned):
    lib = L.lib()
    up = (C.c_float * 3)(*upstream_h)
    gstruct = H.grad_structs(grads) if grads is not None else None
    L.check(lib.kodb_train_tail(
        plan.heads.structs, plan.levels, C.byref(plan.packed.struct), C.byref(plan.cfg),
        plan.pos_weight.data_ptr() if plan.pos_weight is not None else None,
        up, upstream_dev.data_ptr() if upstream_dev is not None else None, int(skip_if_equal),
        gstruct, losses.data_ptr(), n_assigned.data_ptr() if n_assigned is not None else None,
        plan.workspace.data_ptr(), plan.workspace.numel(), H.stream_ptr(plan.heads.device)), "kodb_train_tail")


class _TrainTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: TrainTailPlan, hint, *head_tensors):
        dev = plan.heads.device
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        n_assigned = torch.empty(3, dtype=torch.int32, device=dev)
        need = any(ctx.needs_input_grad[2:])
        grads = [H.alloc_planar_like(t) for t in plan.heads.tensors] if need else None
        launch_train_tail(plan, hint, None, 0, grads, losses, n_assigned)
        ctx.plan, ctx.hint, ctx.grads = plan, hint, grads
        ctx.mark_non_differentiable(n_assigned)
        return losses[0], losses[1], losses[2], n_assigned

    @staticmethod
    def backward(ctx, g_loc, g_obj, g_cls, _g_n):
        plan, grads = ctx.plan, ctx.grads
        dev = plan.heads.device
        gs = []
        for g in (g_loc, g_obj, g_cls):
            if g is None:
                g = torch.zeros((), dtype=torch.float32, device=dev)
            gs.append(g.to(torch.float32))
        up_dev = torch.stack(gs)
        scratch = torch.empty(3, dtype=torch.float32, device=dev)
        launch_train_tail(plan, ctx.hint, up_dev, 1, grads, scratch, None)
        return (None, None) + tuple(grads)


class Yolov5Loss(nn.Module):
    def __init__(self, assigner: Yolov5LabelAssigner, hparams: Yolov5LossParams, iou_calculator: IoUCalculator,
                 weights: Optional[Sequence[float]] = None, upstream_hint: Optional[float] = None):
        super().__init__()
        self.assigner = assigner
        self.hparams = hparams
        self.iou_calculator = iou_calculator
        # loss.py:58-63: a plain attribute (not a buffer); moved to the head's device on first use
        self.weights = torch.tensor([float(w) for w in weights], dtype=torch.float32) if weights is not None else None
        # gradient scale assumed by the fused forward; None -> the batch size (exp.py:124-130)
        self.upstream_hint = upstream_hint
        self.last_num_assigned = None

    def plan(self, image_feature_shape: FeatureShape, net_result, targets) -> TrainTailPlan:
        heads = H.head_structs(net_result)
        dev = heads.device
        hp = self.hparams
        ratios = (hp.lambda_ll_objectness, hp.lambda_ml_objectness, hp.lambda_hl_objectness)
        levels = H.level_cfgs(self.assigner.anchor_info, image_feature_shape.width, image_feature_shape.height, ratios)
        H.check_grid(heads, levels)
        H.check_labels_host(targets, heads.num_classes)
        packed = targets if isinstance(targets, H.PackedTargets) else H.pack_targets(targets, dev)
        if packed.batch != heads.batch:
            raise ValueError(f"{packed.batch} target entries for a batch of {heads.batch}")
        cfg = L.LossCfg()
        cfg.lambda_classification = hp.lambda_classification
        cfg.lambda_localization = hp.lambda_localization
        cfg.lambda_objectness = hp.lambda_objectness
        it = getattr(self.iou_calculator, "iou_type", "ciou")
        cfg.iou_type = L.IOU_TYPES[str(getattr(it, "value", it))]
        cfg.eps = float(getattr(self.iou_calculator, "eps", 1e-7))
        cfg.num_classes = heads.num_classes
        cfg.img_w, cfg.img_h = int(image_feature_shape.width), int(image_feature_shape.height)
        cfg.assign_threshold = float(self.assigner.threshold)
        pw = None
        if self.weights is not None:
            if self.weights.device != dev:
                self.weights = self.weights.to(dev)
            if self.weights.numel() != heads.num_classes:
                raise ValueError("len(weights) must equal the number of classes")
            pw = self.weights
        nbytes = L.lib().kodb_train_workspace_bytes(packed.batch, packed.max_per_image,
                                                    max(lv.num_anchors for lv in levels), heads.num_classes)
        return TrainTailPlan(heads, levels, packed, cfg, pw, H.workspace("train", dev, nbytes))

    def forward(self, image_feature_shape: FeatureShape, net_result, targets) -> LossResult:
        plan = self.plan(image_feature_shape, net_result, targets)
        h = float(self.upstream_hint) if self.upstream_hint is not None else float(plan.heads.batch)
        loc, obj, cls, n_assigned = _TrainTailFn.apply(plan, (h, h, h), *plan.heads.tensors)
        self.last_num_assigned = n_assigned
        return LossResult(localization=loc, objectness=obj, classification=cls)
