"""Oracle (test infrastructure): restatement of the COCO-style matching + AP accumulation that
`PyCOCOMAPEvalCallback` (kod/lightning/callbacks/pycoco_map_eval.py:41-144) obtains from
`vision_evaluation.evaluators.CocoMeanAveragePrecisionEvaluator(ious=[0.3,0.5,0.75,0.9],
report_tag_wise=[False,True,False,False])` -> `pycocotools.cocoeval.COCOeval`.

**PARITY UNPINNED.**  Neither `vision_evaluation` (PyPI `vision-evaluation`, unpinned in the
reference's requirements.txt) nor `pycocotools` (2.0.x) is under /root/reference or installable in
this image, and no reference test holds vectors for this stage.  This file restates the published
pycocotools algorithm (`COCOeval.computeIoU / evaluateImg / accumulate`, `maskApi.c::bbIou`) as
summarised in SURVEY.md Appendix A.7, anchored on the reference's call site:

* inputs (pycoco_map_eval.py:66-104): detections rows (x1,y1,x2,y2,score,cls) -> (label, score,
  l,t,r,b) python floats (fp32 values widened to fp64); GT boxes xyxy + int labels;
* boxes become [x, y, w, h] = [l, t, r-l, b-t] in fp64; no crowd, no ignore, area range "all";
* per (image, class): detections stably sorted by -score (mergesort), at most `max_dets`;
  IoU = i/u with i = w*h when w>0 and h>0 else 0;
* per IoU threshold t: each detection in order takes the un-matched GT with the largest IoU that is
  >= min(t, 1-1e-10); on equal IoU the LATER GT wins (`if iou < best: continue`);
* per class: concatenate (score, tp) over images in image order, stable sort by -score, cumsum,
  precision = tp/(tp+fp+eps64), right-to-left running max, 101-point interpolation with
  searchsorted(side='left'); classes without GT are excluded;
* report keys (pycoco_map_eval.py:115-125): avg_mAP, mAP_30, mAP_50, mAP_75, mAP_90,
  tag_wise_AP_50[class].
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np

IOU_THRESHOLDS = (0.3, 0.5, 0.75, 0.9)
REC_THRS = np.linspace(0.0, 1.0, int(np.round((1.0 - 0.0) / 0.01)) + 1, endpoint=True)


def bbiou(dt_xywh: np.ndarray, gt_xywh: np.ndarray) -> np.ndarray:
    """maskApi.c::bbIou without crowd: (D,4),(G,4) fp64 [x,y,w,h] -> (D,G)."""
    d = np.asarray(dt_xywh, np.float64).reshape(-1, 4)
    g = np.asarray(gt_xywh, np.float64).reshape(-1, 4)
    out = np.zeros((d.shape[0], g.shape[0]), np.float64)
    for gi in range(g.shape[0]):
        ga = g[gi, 2] * g[gi, 3]
        for di in range(d.shape[0]):
            da = d[di, 2] * d[di, 3]
            w = min(d[di, 0] + d[di, 2], g[gi, 0] + g[gi, 2]) - max(d[di, 0], g[gi, 0])
            if w <= 0:
                continue
            h = min(d[di, 1] + d[di, 3], g[gi, 1] + g[gi, 3]) - max(d[di, 1], g[gi, 1])
            if h <= 0:
                continue
            i = w * h
            out[di, gi] = i / (da + ga - i)
    return out


class ImageMatch(NamedTuple):
    score: np.ndarray  # (n,) float32, in the detection order given (NMS order)
    cls: np.ndarray  # (n,) int32
    tp: np.ndarray  # (n, n_thr) bool   tp[d, t]
    gt_count: np.ndarray  # (C,) int64


def match_image(dets: np.ndarray, gt_boxes: np.ndarray, gt_labels: np.ndarray, num_classes: int,
                iou_thrs: Sequence[float] = IOU_THRESHOLDS, max_dets: int = 300) -> ImageMatch:
    """COCOeval.evaluateImg for every class of one image.  dets (n,6) fp32, gt (g,4) xyxy."""
    dets = np.asarray(dets).reshape(-1, 6)
    n = dets.shape[0]
    gtb = np.asarray(gt_boxes, np.float64).reshape(-1, 4)
    gtl = np.asarray(gt_labels).reshape(-1).astype(np.int64)
    d64 = dets.astype(np.float64)
    tp = np.zeros((n, len(iou_thrs)), bool)
    gt_count = np.bincount(gtl, minlength=num_classes).astype(np.int64)[:num_classes]
    dcls = dets[:, 5].astype(np.int64)
    for c in np.unique(dcls):
        didx = np.nonzero(dcls == c)[0]
        didx = didx[np.argsort(-d64[didx, 4], kind="mergesort")][:max_dets]
        gidx = np.nonzero(gtl == c)[0]
        if gidx.size == 0 or didx.size == 0:
            continue
        dxywh = np.stack([d64[didx, 0], d64[didx, 1], d64[didx, 2] - d64[didx, 0], d64[didx, 3] - d64[didx, 1]], 1)
        gxywh = np.stack([gtb[gidx, 0], gtb[gidx, 1], gtb[gidx, 2] - gtb[gidx, 0], gtb[gidx, 3] - gtb[gidx, 1]], 1)
        ious = bbiou(dxywh, gxywh)
        for ti, t in enumerate(iou_thrs):
            taken = np.zeros((gidx.size,), bool)
            for di in range(didx.size):
                best = min(t, 1 - 1e-10)
                m = -1
                for gi in range(gidx.size):
                    if taken[gi]:
                        continue
                    if ious[di, gi] < best:
                        continue
                    best = ious[di, gi]
                    m = gi
                if m >= 0:
                    taken[m] = True
                    tp[didx[di], ti] = True
    return ImageMatch(dets[:, 4].astype(np.float32), dcls.astype(np.int32), tp, gt_count)


def average_precision(scores: np.ndarray, cls: np.ndarray, tp: np.ndarray, gt_count: np.ndarray):
    """COCOeval.accumulate.  scores (N,), cls (N,), tp (N,T) in image-then-detection order.

    Returns ap (T, C) float64 with NaN for classes without GT.
    """
    scores = np.asarray(scores, np.float64).reshape(-1)  # python floats of fp32 values
    cls = np.asarray(cls).reshape(-1)
    tp = np.asarray(tp, bool).reshape(scores.shape[0], -1)
    T = tp.shape[1]
    C = gt_count.shape[0]
    ap = np.full((T, C), np.nan, np.float64)
    for k in range(C):
        npig = int(gt_count[k])
        if npig == 0:
            continue
        sel = np.nonzero(cls == k)[0]
        order = np.argsort(-scores[sel], kind="mergesort")
        tps = tp[sel][order]
        for t in range(T):
            tpc = np.cumsum(tps[:, t]).astype(np.float64)
            fpc = np.cumsum(~tps[:, t]).astype(np.float64)
            nd = tpc.shape[0]
            rc = tpc / npig
            pr = tpc / (fpc + tpc + np.spacing(1))
            q = np.zeros((REC_THRS.shape[0],), np.float64)
            pr = pr.tolist()
            for i in range(nd - 1, 0, -1):
                if pr[i] > pr[i - 1]:
                    pr[i - 1] = pr[i]
            inds = np.searchsorted(rc, REC_THRS, side="left")
            for ri, pi in enumerate(inds):
                if pi >= nd:
                    break
                q[ri] = pr[pi]
            ap[t, k] = float(np.mean(q))
    return ap


def report_from_ap(ap: np.ndarray, iou_thrs: Sequence[float] = IOU_THRESHOLDS) -> dict:
    """vision_evaluation report keys consumed at pycoco_map_eval.py:113-125."""
    rep = {}
    per_t = []
    for ti, t in enumerate(iou_thrs):
        row = ap[ti]
        valid = row[~np.isnan(row)]
        m = float(np.mean(valid)) if valid.size else -1.0
        per_t.append(m)
        rep[f"mAP_{int(round(t * 100))}"] = m
    rep["avg_mAP"] = float(np.mean(per_t)) if per_t else -1.0
    t50 = list(iou_thrs).index(0.5) if 0.5 in iou_thrs else None
    if t50 is not None:
        rep["tag_wise_AP_50"] = {int(k): float(ap[t50, k]) for k in range(ap.shape[1]) if not np.isnan(ap[t50, k])}
    return rep


class CocoMapOracle:
    """`add_predictions / get_report / reset` protocol of the evaluator (pycoco_map_eval.py:106-144)."""

    def __init__(self, num_classes: int, iou_thrs: Sequence[float] = IOU_THRESHOLDS):
        self.num_classes = num_classes
        self.iou_thrs = tuple(iou_thrs)
        self.reset()

    def reset(self):
        self._matches: list[ImageMatch] = []

    def add_image(self, dets, gt_boxes, gt_labels):
        self._matches.append(match_image(dets, gt_boxes, gt_labels, self.num_classes, self.iou_thrs))

    def get_report(self) -> dict:
        if not self._matches:
            ap = np.full((len(self.iou_thrs), self.num_classes), np.nan)
            return report_from_ap(ap, self.iou_thrs)
        s = np.concatenate([m.score for m in self._matches])
        c = np.concatenate([m.cls for m in self._matches])
        tp = np.concatenate([m.tp for m in self._matches], 0)
        g = np.sum([m.gt_count for m in self._matches], 0)
        return report_from_ap(average_precision(s, c, tp, g), self.iou_thrs)
