"""Oracle (test infrastructure): torch-CPU restatement of the YOLOv5 train-tail loss.

Restates `Yolov5Loss.forward` (kod/lightning/experiments/yv5_baseline/loss.py:166-248) with its
three per-level terms (`_compute_localization_loss` :65-98, `_compute_objectness_loss` :100-126,
`_compute_classification_loss` :128-164) and the aligned IoU family of kod/core/bbox/iou.py
(`compute_iou` :77-95, `compute_giou` :136-161, `compute_diou` :164-197, `compute_ciou` :200-246).
Gradients come from torch autograd over this restatement, so the reference's quirks are kept:
the objectness target is NOT detached (obj-BCE back-propagates into the box logits through the
IoU), duplicate cells are written last-writer-wins while every duplicate receives gradient, and an
empty level yields NaN for the loc/cls terms (mean of an empty tensor).

Pinned by tests/golden/loss_*.npz (outputs of the real reference run in the build container).
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from .assign import LevelAssignment


class LossHParams(NamedTuple):
    """`Yolov5LossParams` (loss.py:25-43) with the defaults of configs/nn/losses/yv5.yaml:7-12."""

    lambda_classification: float = 0.5
    lambda_localization: float = 0.05
    lambda_objectness: float = 1.0
    lambda_ll_objectness: float = 4.0
    lambda_ml_objectness: float = 1.0
    lambda_hl_objectness: float = 0.4


def _xyxy(c: torch.Tensor) -> tuple[torch.Tensor, ...]:
    # torchvision _box_cxcywh_to_xyxy: c -/+ 0.5*w
    cx, cy, w, h = c.unbind(-1)
    return cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h


def aligned_iou(pred_cxcywh: torch.Tensor, gt_cxcywh: torch.Tensor, iou_type: str = "ciou", eps: float = 1e-7):
    """iou.py:77-246 on cxcywh boxes (loss.py:84-93 converts before calling the calculator)."""
    return aligned_iou_xyxy(torch.stack(_xyxy(pred_cxcywh), -1), torch.stack(_xyxy(gt_cxcywh), -1), iou_type, eps)


def aligned_iou_xyxy(boxes1: torch.Tensor, boxes2: torch.Tensor, iou_type: str = "ciou", eps: float = 1e-7):
    """IoUCalculator.__call__ (iou.py:249-268) -> compute_{iou,giou,diou,ciou} (iou.py:77-246) on (..., 4) xyxy
    boxes.  Index 1 = boxes1 (prediction), g = boxes2 (ground truth).  Pinned by tests/golden/iou_*.npz."""
    x1, y1, x2, y2 = boxes1.unbind(-1)
    x1g, y1g, x2g, y2g = boxes2.unbind(-1)
    iw = (torch.min(x2, x2g) - torch.max(x1, x1g)).clamp(0)
    ih = (torch.min(y2, y2g) - torch.max(y1, y1g)).clamp(0)
    inter = iw * ih
    union = (x2 - x1) * (y2 - y1) + (x2g - x1g) * (y2g - y1g) - inter
    iou = inter / (union + eps)
    if iou_type == "iou":
        return iou
    cw = torch.max(x2, x2g) - torch.min(x1, x1g)
    ch = torch.max(y2, y2g) - torch.min(y1, y1g)
    if iou_type == "giou":
        hull = cw * ch
        return iou - torch.abs(hull - union) / torch.abs(hull + eps)
    diag2 = cw**2 + ch**2
    cdist2 = ((x1 + x2) / 2 - (x1g + x2g) / 2) ** 2 + ((y1 + y2) / 2 - (y1g + y2g) / 2) ** 2
    D = cdist2 / (diag2 + eps)
    if iou_type == "diou":
        return iou - D
    assert iou_type == "ciou", iou_type
    w1, h1 = x2 - x1, y2 - y1
    w2, h2 = x2g - x1g, y2g - y1g
    v = (4 / math.pi**2) * torch.pow(torch.atan(w2 / (h2 + eps)) - torch.atan(w1 / (h1 + eps)), 2)
    with torch.no_grad():
        alpha = v / ((1 - iou) + v + eps)
    return iou - D - alpha * v


def _as_idx(a: LevelAssignment):
    f = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.long)
    return f(a.samples), f(a.anchors), f(a.grid_y), f(a.grid_x)


def level_terms(
    box: torch.Tensor,  # (B,A,H,W,4) logits
    obj: torch.Tensor,  # (B,A,H,W,1)
    cls: torch.Tensor,  # (B,A,H,W,C)
    a: LevelAssignment,
    obj_ratio: float,
    iou_type: str = "ciou",
    eps: float = 1e-7,
    pos_weight: Optional[torch.Tensor] = None,
):
    idx = _as_idx(a)
    dt = box.dtype
    gt = torch.as_tensor(a.gt_boxes).to(dt)
    anc = torch.as_tensor(a.anchor_wh).to(dt)
    labels = torch.as_tensor(a.labels, dtype=torch.long)

    p = box[idx]  # (N,4)
    pxy = p[:, :2].sigmoid() * 2 - 0.5
    pwh = ((p[:, 2:4].sigmoid() * 2) ** 2) * anc
    iou = aligned_iou(torch.cat((pxy, pwh), 1), gt, iou_type, eps).squeeze()
    loc = (1 - iou).mean()

    tobj = torch.zeros_like(obj).squeeze(-1)
    tobj[idx] = iou.clamp(0).type(tobj.dtype)  # plain index_put_: last writer wins on CPU
    objl = obj_ratio * F.binary_cross_entropy_with_logits(obj, tobj.unsqueeze(-1), reduction="mean")

    q = cls[idx]  # (N,C)
    onehot = torch.zeros_like(q)
    onehot[torch.arange(q.shape[0]), labels] = 1
    clsl = F.binary_cross_entropy_with_logits(q, onehot, reduction="mean", pos_weight=pos_weight)
    return loc, objl, clsl, iou


def loss_forward(
    heads: Sequence[tuple[torch.Tensor, torch.Tensor, torch.Tensor]],  # [(box,obj,cls)] for ll, ml, hl
    assignment: Sequence[LevelAssignment],
    img_w: int,
    hp: LossHParams = LossHParams(),
    iou_type: str = "ciou",
    eps: float = 1e-7,
    weights: Optional[Sequence[float]] = None,
):
    """Returns (localization, objectness, classification) scaled exactly like loss.py:225-248."""
    pw = None
    if weights is not None:
        pw = torch.as_tensor(np.asarray(weights, dtype=np.float32)).to(heads[0][0].dtype)
    ratios = (hp.lambda_ll_objectness, hp.lambda_ml_objectness, hp.lambda_hl_objectness)
    loc = obj = cls = None
    for (b, o, c), a, r in zip(heads, assignment, ratios):
        l, ob, cl, _ = level_terms(b, o, c, a, r, iou_type, eps, pw)
        loc = l if loc is None else loc + l
        obj = ob if obj is None else obj + ob
        cls = cl if cls is None else cls + cl
    num_classes = heads[0][2].shape[-1]
    lam_obj = hp.lambda_objectness * ((img_w / 640) ** 2)
    lam_cls = hp.lambda_classification * (num_classes / 80)
    return hp.lambda_localization * loc, lam_obj * obj, lam_cls * cls


def loss_forward_backward(heads, assignment, img_w, hp=LossHParams(), iou_type="ciou", eps=1e-7,
                          weights=None, upstream=(1.0, 1.0, 1.0)):
    """Loss triple + gradients of `u0*loc + u1*obj + u2*cls` w.r.t. the nine head tensors.

    `training_step` (exp.py:124-130) uses upstream = (B, B, B).  Gradients are returned in the same
    (B,A,H,W,P) logical shape as the inputs.
    """
    leaves = []
    for (b, o, c) in heads:
        leaves.append(tuple(t.detach().clone().requires_grad_(True) for t in (b, o, c)))
    loc, obj, cls = loss_forward(leaves, assignment, img_w, hp, iou_type, eps, weights)
    total = upstream[0] * loc + upstream[1] * obj + upstream[2] * cls
    flat = [t for lv in leaves for t in lv]
    grads = torch.autograd.grad(total, flat, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(t) for g, t in zip(grads, flat)]
    grads = [tuple(grads[3 * i : 3 * i + 3]) for i in range(len(leaves))]
    return (loc.detach(), obj.detach(), cls.detach()), grads


def class_weights_from_counts(counts: Sequence[float]) -> np.ndarray:
    """`sum(n)/n_c` in float32 (kod/lightning/tasks/trainer.py:54-60)."""
    w = np.array(list(counts), dtype=np.float32)
    return (np.sum(w) / w).astype(np.float32)
