"""`IoUType` / `IoUCalculator` (kod/core/bbox/iou.py:9-14, :249-268).

`IoUCalculator.__call__(boxes1, boxes2)` evaluates the aligned IoU family (compute_iou / giou / diou / ciou,
iou.py:77-246) on two `(..., 4)` xyxy tensors with the CUDA kernel `kodb_iou_aligned` (csrc/iou.cu), forward and
backward (autograd to both inputs, CIoU's alpha treated as a constant like the reference's `no_grad` block).  Inside
`Yolov5Loss` the calculator only carries `(iou_type, eps)` into the fused train tail (csrc/train_tail.cu::iou_metric),
which evaluates the same expression on the decoded head logits.
"""
from __future__ import annotations

import enum

import torch

from . import _host as H
from . import _lib as L


@enum.unique
class IoUType(str, enum.Enum):
    ioU = "iou"
    giou = "giou"
    diou = "diou"
    ciou = "ciou"


class _IoUAlignedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes1, boxes2, iou_type: int, eps: float):
        b1 = boxes1.detach().contiguous()
        b2 = boxes2.detach().contiguous()
        n = b1.numel() // 4
        out = torch.empty(boxes1.shape[:-1], dtype=torch.float32, device=b1.device)
        L.check(L.lib().kodb_iou_aligned(b1.data_ptr(), b2.data_ptr(), n, iou_type, eps, out.data_ptr(), None, None,
                                         None, H.stream_ptr(b1.device)), "kodb_iou_aligned")
        ctx.save_for_backward(b1, b2)
        ctx.cfg = (iou_type, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        b1, b2 = ctx.saved_tensors
        iou_type, eps = ctx.cfg
        n = b1.numel() // 4
        up = g.contiguous().to(torch.float32)
        g1 = torch.empty_like(b1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(b2) if ctx.needs_input_grad[1] else None
        if g1 is not None or g2 is not None:
            L.check(L.lib().kodb_iou_aligned(b1.data_ptr(), b2.data_ptr(), n, iou_type, eps, None, up.data_ptr(),
                                             g1.data_ptr() if g1 is not None else None,
                                             g2.data_ptr() if g2 is not None else None, H.stream_ptr(b1.device)),
                    "kodb_iou_aligned")
        return g1, g2, None, None


class IoUCalculator(object):
    def __init__(self, iou_type: IoUType = IoUType.ciou, eps: float = 1e-7):
        self.iou_type = IoUType(iou_type)
        self.eps = float(eps)

    def __call__(self, boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
        assert boxes1.shape == boxes2.shape  # iou.py:82
        if boxes1.shape[-1] != 4:
            raise ValueError("boxes must be (..., 4) xyxy")
        H.require_cuda(boxes1, "boxes1")
        H.require_cuda(boxes2, "boxes2")
        if boxes1.dtype != torch.float32 or boxes2.dtype != torch.float32:
            raise TypeError("boxes must be float32 (the reference trains in fp32)")
        return _IoUAlignedFn.apply(boxes1, boxes2, L.IOU_TYPES[self.iou_type.value], self.eps)
