"""Eval-tail drop-ins on the CUDA kernels of csrc/decode.cu and csrc/eval.cu.

* `get_detections(image_feature_shape, net_result, anchor_info)` – DefaultYolov5Experiment.get_detections
  (kod/lightning/experiments/yv5_baseline/exp.py:70-102): (B, AC, 5+C).  The head tensors are left
  untouched (the reference overwrites the box head in place, layers.py:57-60; callers never read it again).
* `non_max_suppression(detections, conf_thres, nms_thres, classes)` – kod/core/nms.py:9-75, same return
  value: a list of B tensors (n_i, 6).  Exact lengths need the per-image counts on the host: one sync.
* `nms_padded(...)` / `eval_tail(...)` – the no-sync forms: padded (B, 300, 6) detections, int64 keep
  indices into the candidate list, device counts; `eval_tail` fuses decode + NMS (+ COCO matching) and
  never materialises the (B, AC, 5+C) tensor.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import torch

from . import _host as H
from . import _lib as L
from .types import FeatureShape

MAX_WH = 4096  # kod/core/nms.py:24-27
MAX_DET = 300
MAX_NMS = 30000
IOU_THRESHOLDS = (0.3, 0.5, 0.75, 0.9)  # pycoco_map_eval.py:45-48


class NmsResult(NamedTuple):
    dets: torch.Tensor  # (B, max_det, 6) x1,y1,x2,y2,conf,cls; rows >= counts[b] are zero
    keep_idx: torch.Tensor  # (B, max_det) int64 index into the candidate list handed to NMS; -1 padding
    counts: torch.Tensor  # (B,) int32
    num_candidates: torch.Tensor  # (B,) int32 candidates before the max_nms cut
    keep_key: torch.Tensor = None  # (B, max_det) int64 row * C + class of every kept detection; -1 padding


class MatchRecords(NamedTuple):
    score: torch.Tensor  # (B, max_det) float32
    cls: torch.Tensor  # (B, max_det) int32
    tp: torch.Tensor  # (B, max_det) uint8, bit t = TP at IOU_THRESHOLDS[t]
    counts: torch.Tensor  # (B,) int32


def get_detections(image_feature_shape: FeatureShape, net_result, anchor_info) -> torch.Tensor:
    heads = H.head_structs(net_result)
    levels = H.level_cfgs(anchor_info, image_feature_shape.width, image_feature_shape.height)
    H.check_grid(heads, levels)
    rows = sum(lv.num_anchors * lv.H * lv.W for lv in levels)
    out = torch.empty((heads.batch, rows, 5 + heads.num_classes), dtype=torch.float32, device=heads.device)
    L.check(L.lib().kodb_decode(heads.structs, levels, heads.batch, heads.num_classes, out.data_ptr(),
                                H.stream_ptr(heads.device)), "kodb_decode")
    return out


def _nms_cfg(num_classes: int, conf_thres: float, nms_thres: float, classes, max_det: int, max_nms: int) -> L.NmsCfg:
    cfg = L.NmsCfg()
    cfg.conf_thres = float(conf_thres)
    cfg.iou_thres = float(nms_thres)
    cfg.max_wh = float(MAX_WH)
    cfg.max_det = int(max_det)
    cfg.max_nms = int(max_nms)
    cfg.num_classes = int(num_classes)
    words = [0] * (L.MAX_CLASSES // 32)
    allowed = range(num_classes) if classes is None else [int(c) for c in classes]
    for c in allowed:
        if 0 <= c < L.MAX_CLASSES:
            words[c // 32] |= 1 << (c % 32)
    for i, w in enumerate(words):
        cfg.class_mask[i] = w
    return cfg


def _alloc_out(B: int, max_det: int, device):
    dets = torch.empty((B, max_det, 6), dtype=torch.float32, device=device)
    keep = torch.empty((B, max_det), dtype=torch.long, device=device)
    counts = torch.empty((B,), dtype=torch.int32, device=device)
    ncand = torch.empty((B,), dtype=torch.int32, device=device)
    key = torch.empty((B, max_det), dtype=torch.long, device=device)
    out = L.NmsOut()
    out.dets, out.keep_idx, out.counts, out.num_candidates = (dets.data_ptr(), keep.data_ptr(), counts.data_ptr(),
                                                              ncand.data_ptr())
    out.keep_key = key.data_ptr()
    return NmsResult(dets, keep, counts, ncand, key), out


class MatchState(NamedTuple):
    """Device buffers the matching stage writes for one batch (+ the running per-class GT count)."""

    records: MatchRecords
    struct: L.MatchOut


def _alloc_match(B: int, max_det: int, gt_count: torch.Tensor, counts: torch.Tensor, device,
                 iou_thrs: Sequence[float]) -> MatchState:
    score = torch.empty((B, max_det), dtype=torch.float32, device=device)
    cls = torch.empty((B, max_det), dtype=torch.int32, device=device)
    tp = torch.empty((B, max_det), dtype=torch.uint8, device=device)
    m = L.MatchOut()
    m.rec_score, m.rec_cls, m.rec_tp, m.gt_count = score.data_ptr(), cls.data_ptr(), tp.data_ptr(), gt_count.data_ptr()
    if len(iou_thrs) != L.NUM_IOU_THRS:
        raise ValueError(f"exactly {L.NUM_IOU_THRS} IoU thresholds are supported (pycoco_map_eval.py:45-48)")
    for i, t in enumerate(iou_thrs):
        m.iou_thrs[i] = float(t)
    return MatchState(MatchRecords(score, cls, tp, counts), m)


def nms_padded(detections: torch.Tensor, conf_thres: float = 0.25, nms_thres: float = 0.45, classes=None,
               max_det: int = MAX_DET, max_nms: int = MAX_NMS, targets=None, gt_count: Optional[torch.Tensor] = None,
               iou_thrs: Sequence[float] = IOU_THRESHOLDS):
    """kodb_nms on a decoded (B, N, 5+C) tensor; with `targets` also the COCO matching (gt_count += )."""
    H.require_cuda(detections, "detections")
    if detections.dtype != torch.float32:
        raise TypeError("detections must be float32")
    det = detections.contiguous()
    B, N, cols = det.shape
    C = cols - 5
    dev = det.device
    lib = L.lib()
    cfg = _nms_cfg(C, conf_thres, nms_thres, classes, max_det, max_nms)
    res, out = _alloc_out(B, max_det, dev)
    ws = H.workspace("eval", dev, lib.kodb_eval_workspace_bytes(B, N, C), tag=B)
    packed, match = None, None
    if targets is not None:
        packed = targets if isinstance(targets, H.PackedTargets) else H.pack_targets(targets, dev)
        match = _alloc_match(B, max_det, gt_count, res.counts, dev, iou_thrs)
    L.check(lib.kodb_nms(det.data_ptr(), B, N, cfg, out, packed.struct if packed else None,
                         match.struct if match else None, ws.data_ptr(), ws.numel(), H.stream_ptr(dev)), "kodb_nms")
    return (res, match.records) if match else res


def eval_tail(image_feature_shape: FeatureShape, net_result, anchor_info, conf_thres: float = 0.001,
              nms_thres: float = 0.6, classes=None, max_det: int = MAX_DET, max_nms: int = MAX_NMS, targets=None,
              gt_count: Optional[torch.Tensor] = None, iou_thrs: Sequence[float] = IOU_THRESHOLDS):
    """validation_step's decode + NMS (exp.py:140-154) fused; optionally + the COCO matching."""
    heads = H.head_structs(net_result)
    levels = H.level_cfgs(anchor_info, image_feature_shape.width, image_feature_shape.height)
    H.check_grid(heads, levels)
    dev = heads.device
    lib = L.lib()
    rows = sum(lv.num_anchors * lv.H * lv.W for lv in levels)
    cfg = _nms_cfg(heads.num_classes, conf_thres, nms_thres, classes, max_det, max_nms)
    res, out = _alloc_out(heads.batch, max_det, dev)
    ws = H.workspace("eval", dev, lib.kodb_eval_workspace_bytes(heads.batch, rows, heads.num_classes), tag=heads.batch)
    packed, match = None, None
    if targets is not None:
        packed = targets if isinstance(targets, H.PackedTargets) else H.pack_targets(targets, dev)
        match = _alloc_match(heads.batch, max_det, gt_count, res.counts, dev, iou_thrs)
    L.check(lib.kodb_eval_tail(heads.structs, levels, heads.batch, cfg, out, packed.struct if packed else None,
                               match.struct if match else None, ws.data_ptr(), ws.numel(), H.stream_ptr(dev)),
            "kodb_eval_tail")
    return (res, match.records) if match else res


def unpad(res: NmsResult) -> list:
    """Padded result -> the reference's list of (n_i, 6) tensors (one host sync for the counts)."""
    counts = res.counts.tolist()
    return [res.dets[b, :n] for b, n in enumerate(counts)]


def non_max_suppression(detections: torch.Tensor, conf_thres: float = 0.25, nms_thres: float = 0.45,
                        classes=None) -> Sequence[torch.Tensor]:
    return unpad(nms_padded(detections, conf_thres, nms_thres, classes))
