"""Drop-in `Yolov5Head` (kod/nn/heads/yolov5.py:134-178) on the tcgen05 kernel `kodb_head_forward` (csrc/head.cu):
the box / objectness / class 1x1 convolutions of a pyramid level as ONE tf32 tensor-core GEMM that reads the feature map
once and writes the three planar `[B][A*P][H][W]` tensors the detection tail consumes.

* `Yolov5Head(in_channels, num_anchors_per_cell, num_classes, stride, prior_probability, use_yv5_init)` – same constructor,
  same sub-module / parameter names (`box_head.conv.weight`, ...: state dicts are interchangeable with the reference),
  same bias initialisation (yolov5.py:62-75, :111-123), same output `DetectionHeadResult(box, obj, cls)` of
  `(B, A, H, W, P)` views.
* `fused_heads_forward(heads, xs)` – all (up to three) levels in ONE launch.
* `fused_heads_detections(heads, xs, image_feature_shape, anchor_info)` – the eval side: `get_detections` (exp.py:70-102) of
  those head outputs in the SAME launch – the GEMM's epilogue decodes the accumulators (sigmoid / grid / anchor) and writes
  the `(B, AC, 5+C)` rows; the planar logits are never written.  No gradient (inference / validation).
* Backward: dX / dW / db with library GEMMs (cuBLAS through torch) – the gradients of the planar tensors come straight
  from the train tail, already in the layout the GEMMs want.
Precision: fp32 in, tf32 multiply, fp32 accumulate, i.e. what torch's default convolution policy
(`torch.backends.cudnn.allow_tf32 = True`) computes on this GPU.  Levels outside the tensor-core path (in_channels % 32,
H*W % 4 – the 13x13 grid of a 416 input –, A*(5+C) > 256) run on a SIMT kernel of the same library in fp32; there is no
torch / cuDNN fallback.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn as nn

from . import _host as H
from . import _lib as L
from .types import DetectionHeadResult


class _Conv1x1Head(nn.Module):
    """Holds `self.conv` like the reference's Yolov5BoxHead / ObjectnessHead / ClassificationHead."""

    def __init__(self, in_channels: int, out_channels: int, bias_shift: float = 0.0, num_anchors: int = 1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1)
        if bias_shift != 0.0:
            with torch.no_grad():
                self.conv.bias.view(num_anchors, -1).add_(bias_shift)


class _FusedHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, num_anchors: int, num_classes: int, *tensors):
        """tensors: per level (x, w_box, w_obj, w_cls, b_box, b_obj, b_cls)."""
        n_lv = len(tensors) // 7
        lib = L.lib()
        arr = (L.HeadConvLevel * n_lv)()
        outs, keep = [], []
        B = tensors[0].shape[0]
        dev = tensors[0].device
        for l in range(n_lv):
            x, wb, wo, wc, bb, bo, bc = tensors[7 * l:7 * l + 7]
            H.require_cuda(x, "feature map")
            if x.dtype != torch.float32:
                raise TypeError("the head runs in float32 (tf32 tensor cores), like the reference")
            x = x.contiguous()
            K, Hh, Ww = x.shape[1], x.shape[2], x.shape[3]
            ws = [w.detach().reshape(w.shape[0], K) for w in (wb, wo, wc)]
            # the three weights back to back without a copy when they already are (one arena / one flat buffer)
            n = [w.shape[0] for w in ws]
            base = ws[0].data_ptr()
            if (all(w.is_contiguous() for w in ws) and ws[1].data_ptr() == base + n[0] * K * 4
                    and ws[2].data_ptr() == base + (n[0] + n[1]) * K * 4 and base % 16 == 0):
                w_all = ws[0]
            else:
                w_all = torch.cat(ws, 0)
            ob = torch.empty((B, n[0], Hh, Ww), dtype=torch.float32, device=dev)
            oo = torch.empty((B, n[1], Hh, Ww), dtype=torch.float32, device=dev)
            oc = torch.empty((B, n[2], Hh, Ww), dtype=torch.float32, device=dev)
            a = arr[l]
            a.x, a.weight = x.data_ptr(), w_all.data_ptr()
            a.bias_box, a.bias_obj, a.bias_cls = bb.data_ptr(), bo.data_ptr(), bc.data_ptr()
            a.out_box, a.out_obj, a.out_cls = ob.data_ptr(), oo.data_ptr(), oc.data_ptr()
            a.in_channels, a.H, a.W = K, Hh, Ww
            outs += [ob, oo, oc]
            keep += [x, w_all]
        L.check(lib.kodb_head_forward(arr, n_lv, B, num_anchors, num_classes, H.stream_ptr(dev)), "kodb_head_forward")
        ctx.save_for_backward(*tensors)
        ctx.n_lv = n_lv
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        tensors = ctx.saved_tensors
        res = [None, None]
        for l in range(ctx.n_lv):
            x, wb, wo, wc, bb, bo, bc = tensors[7 * l:7 * l + 7]
            B, K = x.shape[0], x.shape[1]
            xf = x.reshape(B, K, -1)  # (B, K, HW)
            gs = [g.reshape(B, g.shape[1], -1) if g is not None else None for g in grads[3 * l:3 * l + 3]]
            need = ctx.needs_input_grad[2 + 7 * l:2 + 7 * l + 7]
            dx = None
            if need[0]:
                for w, g in zip((wb, wo, wc), gs):
                    if g is None:
                        continue
                    wt = w.reshape(w.shape[0], K).t()  # (K, n)
                    dx = torch.matmul(wt, g) if dx is None else dx.baddbmm_(wt.unsqueeze(0).expand(B, -1, -1), g)
                dx = dx.reshape(x.shape) if dx is not None else None
            dws = [torch.einsum("bnp,bkp->nk", g, xf).reshape(w.shape) if (nd and g is not None) else None
                   for w, g, nd in zip((wb, wo, wc), gs, need[1:4])]
            dbs = [g.sum((0, 2)) if (nd and g is not None) else None for g, nd in zip(gs, need[4:7])]
            res += [dx] + dws + dbs
        return tuple(res)


def fused_heads_forward(heads: Sequence["Yolov5Head"], xs: Sequence[torch.Tensor]) -> list:
    """`[head(x) for head, x in zip(heads, xs)]` in one launch (1..3 levels, same A and C)."""
    A, C = heads[0].num_anchors, heads[0].num_classes
    args = []
    for h, x in zip(heads, xs):
        if (h.num_anchors, h.num_classes) != (A, C):
            raise ValueError("levels of one launch share the anchor and class counts")
        args += [x, h.box_head.conv.weight, h.obj_head.conv.weight, h.cls_head.conv.weight, h.box_head.conv.bias,
                 h.obj_head.conv.bias, h.cls_head.conv.bias]
    outs = _FusedHeadsFn.apply(A, C, *args)
    res = []
    for l in range(len(heads)):
        planes = outs[3 * l:3 * l + 3]
        # Rearrange("b (a p) h w -> b a h w p") as a view (yolov5.py:29-32, :77-80, :125-128)
        res.append(DetectionHeadResult(*[t.view(t.shape[0], A, t.shape[1] // A, t.shape[2], t.shape[3]).permute(0, 1, 3, 4, 2)
                                         for t in planes]))
    return res


def _level_args(heads, xs):
    """ctypes level array + the tensors that must stay alive during the launch."""
    n_lv = len(heads)
    arr = (L.HeadConvLevel * n_lv)()
    keep = []
    for l, (h, x) in enumerate(zip(heads, xs)):
        H.require_cuda(x, "feature map")
        if x.dtype != torch.float32:
            raise TypeError("the head runs in float32 (tf32 tensor cores), like the reference")
        x = x.detach().contiguous()
        K = x.shape[1]
        ws = [m.conv.weight.detach().reshape(m.conv.weight.shape[0], K) for m in (h.box_head, h.obj_head, h.cls_head)]
        n = [w.shape[0] for w in ws]
        base = ws[0].data_ptr()
        if (all(w.is_contiguous() for w in ws) and ws[1].data_ptr() == base + n[0] * K * 4
                and ws[2].data_ptr() == base + (n[0] + n[1]) * K * 4 and base % 16 == 0):
            w_all = ws[0]
        else:
            w_all = torch.cat(ws, 0)
        a = arr[l]
        a.x, a.weight = x.data_ptr(), w_all.data_ptr()
        a.bias_box, a.bias_obj, a.bias_cls = (m.conv.bias.data_ptr() for m in (h.box_head, h.obj_head, h.cls_head))
        a.in_channels, a.H, a.W = K, x.shape[2], x.shape[3]
        keep += [x, w_all]
    return arr, keep


def fused_heads_detections(heads: Sequence["Yolov5Head"], xs: Sequence[torch.Tensor], image_feature_shape, anchor_info) -> torch.Tensor:
    """`get_detections(image_feature_shape, [head(x) for head, x in zip(heads, xs)])` (exp.py:70-102) in ONE launch: the
    `(B, AC, 5+C)` detections straight from the feature maps.  Levels outside the tensor-core path (see the module
    docstring) take two launches instead: planar logits, then the decode kernel."""
    from . import nms as N
    A, C = heads[0].num_anchors, heads[0].num_classes
    levels = H.level_cfgs(anchor_info, image_feature_shape.width, image_feature_shape.height)
    B, dev = xs[0].shape[0], xs[0].device
    tc = all(x.shape[1] % 32 == 0 and (x.shape[2] * x.shape[3]) % 4 == 0 for x in xs) and A * (5 + C) <= 256
    if not tc or len(heads) != L.NUM_LEVELS:
        with torch.no_grad():
            return N.get_detections(image_feature_shape, fused_heads_forward(heads, xs), anchor_info)
    for lv, x in zip(levels, xs):
        if (lv.H, lv.W) != (x.shape[2], x.shape[3]) or lv.num_anchors != A:
            raise ValueError("feature maps do not match image_feature_shape // stride (layers.py:24-27)")
    arr, keep = _level_args(heads, xs)
    rows = sum(A * x.shape[2] * x.shape[3] for x in xs)
    out = torch.empty((B, rows, 5 + C), dtype=torch.float32, device=dev)
    L.check(L.lib().kodb_head_detections(arr, levels, len(heads), B, A, C, out.data_ptr(), H.stream_ptr(dev)),
            "kodb_head_detections")
    return out


class Yolov5Head(nn.Module):
    def __init__(self, in_channels: int, num_anchors_per_cell: int, num_classes: int, stride: int,
                 prior_probability: float = 0.01, use_yv5_init: bool = True):
        super().__init__()
        self.num_anchors, self.num_classes = int(num_anchors_per_cell), int(num_classes)
        prior = -math.log((1 - prior_probability) / prior_probability)
        obj_shift = math.log(8 / (640 / stride) ** 2) if use_yv5_init else prior        # yolov5.py:113-116
        cls_shift = math.log(0.6 / (num_classes - 0.99999)) if use_yv5_init else prior   # yolov5.py:65-68
        A = self.num_anchors
        self.box_head = _Conv1x1Head(in_channels, A * 4)
        self.obj_head = _Conv1x1Head(in_channels, A, obj_shift, A)
        self.cls_head = _Conv1x1Head(in_channels, A * num_classes, cls_shift, A)
        # the three weights as views of one flat buffer: the kernel reads them back to back without a copy (as long as
        # nobody re-homes the parameters one by one; an arena optimizer keeps them adjacent, anything else costs a cat)
        K = in_channels
        flat = torch.cat([m.conv.weight.detach().reshape(-1) for m in (self.box_head, self.obj_head, self.cls_head)])
        o = 0
        for m in (self.box_head, self.obj_head, self.cls_head):
            n = m.conv.weight.shape[0]
            m.conv.weight = nn.Parameter(flat[o:o + n * K].view(n, K, 1, 1))
            o += n * K

    def forward(self, x: torch.Tensor) -> DetectionHeadResult:
        return fused_heads_forward([self], [x])[0]
