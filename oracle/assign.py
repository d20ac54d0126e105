"""Oracle (test infrastructure): numpy restatement of the YOLOv5 anchor/grid target assignment.

Follows `Yolov5LabelAssigner._assign_per_layer` (kod/core/label_assignment/yv5.py:207-296) and its
helpers `_make_batched_targets` (:68-121), `_filter_targets` (:160-176),
`_incorporate_neighbouring_cells` (:178-205); box conversion is torchvision's
`_box_xyxy_to_cxcywh` ((x1+x2)/2, x2-x1); scaling is `CXCYWHBoundingBox.scale_tensor`
(kod/core/bbox/boxes.py:45-55).

All compares are done in float32 exactly like the reference's fp32 row buffer; the integer outputs
are therefore bit-exact restatements (pinned by tests/golden/assign_*.npz).
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np

# (dx, dy) * 0.5 in the reference's order: base cell, left, up, right, down  (yv5.py:56-66)
NEIGHBOUR_OFFSETS = np.array(
    [[0.0, 0.0], [0.5, 0.0], [0.0, 0.5], [-0.5, 0.0], [0.0, -0.5]], dtype=np.float32
)


class LevelAssignment(NamedTuple):
    samples: np.ndarray  # (N,) int64
    anchors: np.ndarray  # (N,) int64
    grid_y: np.ndarray  # (N,) int64
    grid_x: np.ndarray  # (N,) int64
    labels: np.ndarray  # (N,) int64
    gt_boxes: np.ndarray  # (N,4) float32  (cx-gi, cy-gj, w, h) in grid units
    anchor_wh: np.ndarray  # (N,2) float32  anchor / stride
    # provenance of every row (not part of the reference output; used by tests of the fused kernel)
    src_target: np.ndarray  # (N,) int64  index into the flat target list
    src_offset: np.ndarray  # (N,) int64  0..4


def flatten_targets(targets: Sequence) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """tuple of per-image (boxes, labels) -> flat boxes (T,4), labels (T,), image offsets (B+1,)."""
    boxes, labels, offs = [], [], [0]
    dtype = None
    for t in targets:
        b = np.asarray(t[0])
        l = np.asarray(t[1])
        b = b.reshape(-1, 4) if b.size else np.zeros((0, 4), b.dtype if b.dtype.kind == "f" else np.float32)
        if b.shape[0] and dtype is None:
            dtype = b.dtype
        boxes.append(b)
        labels.append(l.reshape(-1).astype(np.int64))
        offs.append(offs[-1] + b.shape[0])
    dtype = dtype or np.float32
    fb = np.concatenate([b.astype(dtype) for b in boxes], 0) if boxes else np.zeros((0, 4), dtype)
    fl = np.concatenate(labels, 0) if labels else np.zeros((0,), np.int64)
    return fb, fl, np.asarray(offs, dtype=np.int32)


def scaled_anchors(anchors_px: Sequence[Sequence[float]], stride: int) -> np.ndarray:
    """`a.width * 1 / stride` in python floats, then rounded to fp32 (yv5.py:227-239)."""
    return np.array([[a[0] * 1 / stride, a[1] * 1 / stride] for a in anchors_px], dtype=np.float32)


def _py_remainder_one(x: np.ndarray) -> np.ndarray:
    """torch.remainder(x, 1) for floats: fmod, then shift negative results by +1."""
    r = np.fmod(x, x.dtype.type(1))
    neg = r < 0
    r = np.where(neg, r + x.dtype.type(1), r)
    return r.astype(x.dtype)


def assign_level(
    boxes_xyxy: np.ndarray,
    labels: np.ndarray,
    img_offsets: np.ndarray,
    stride: int,
    anchors_px: Sequence[Sequence[float]],
    img_w: int,
    img_h: int,
    threshold: float = 4.0,
) -> LevelAssignment:
    T = int(boxes_xyxy.shape[0])
    B = len(img_offsets) - 1
    img_ids = np.zeros((T,), np.int64)
    for b in range(B):
        img_ids[img_offsets[b] : img_offsets[b + 1]] = b

    # -- step 1: xyxy -> cxcywh in the input dtype, scale by 1/stride, store as fp32 ------------
    bx = np.asarray(boxes_xyxy)
    in_t = bx.dtype.type
    scale = 1 / stride
    if T:
        cx = ((bx[:, 0] + bx[:, 2]) / in_t(2)) * in_t(scale)
        cy = ((bx[:, 1] + bx[:, 3]) / in_t(2)) * in_t(scale)
        w = (bx[:, 2] - bx[:, 0]) * in_t(scale)
        h = (bx[:, 3] - bx[:, 1]) * in_t(scale)
    else:
        cx = cy = w = h = np.zeros((0,), bx.dtype)
    cx, cy, w, h = (v.astype(np.float32) for v in (cx, cy, w, h))
    lab_f = labels.astype(np.float32)  # the reference stores the label in the fp32 row

    anc = scaled_anchors(anchors_px, stride)  # (A,2) fp32
    A = anc.shape[0]

    # -- step 2: anchor filter, anchor-major then target order ----------------------------------
    with np.errstate(divide="ignore", invalid="ignore"):
        rw = w[None, :] / anc[:, 0:1]
        rh = h[None, :] / anc[:, 1:2]
        one = np.float32(1.0)
        mw = np.maximum(rw, one / rw)
        mh = np.maximum(rh, one / rh)
        keep = np.maximum(mw, mh) < np.float32(threshold)  # (A,T)
    k_idx, n_idx = np.nonzero(keep)  # row-major: anchor-major, then n

    # -- step 3: neighbour-cell expansion --------------------------------------------------------
    fm_w = np.float32(img_w / stride)
    fm_h = np.float32(img_h / stride)
    gx, gy = cx[n_idx], cy[n_idx]
    ix, iy = fm_w - gx, fm_h - gy
    half = np.float32(0.5)
    j = (_py_remainder_one(gx) < half) & (gx > one)
    k = (_py_remainder_one(gy) < half) & (gy > one)
    l = (_py_remainder_one(ix) < half) & (ix > one)
    m = (_py_remainder_one(iy) < half) & (iy > one)
    masks = np.stack([np.ones_like(j), j, k, l, m])  # (5,M)
    o_idx, r_idx = np.nonzero(masks)  # offset-major, then filtered-row order

    n_sel = n_idx[r_idx]
    k_sel = k_idx[r_idx]
    off = NEIGHBOUR_OFFSETS[o_idx]  # (N,2)
    cxs, cys = cx[n_sel], cy[n_sel]
    gi = np.trunc(cxs - off[:, 0]).astype(np.int64)
    gj = np.trunc(cys - off[:, 1]).astype(np.int64)
    W = img_w // stride
    H = img_h // stride
    gt = np.stack(
        [cxs - gi.astype(np.float32), cys - gj.astype(np.float32), w[n_sel], h[n_sel]], axis=1
    ).astype(np.float32)
    return LevelAssignment(
        samples=img_ids[n_sel].astype(np.int64),
        anchors=k_sel.astype(np.int64),
        grid_y=np.clip(gj, 0, H - 1).astype(np.int64),
        grid_x=np.clip(gi, 0, W - 1).astype(np.int64),
        labels=lab_f[n_sel].astype(np.int64),
        gt_boxes=gt.reshape(-1, 4),
        anchor_wh=anc[k_sel].reshape(-1, 2),
        src_target=n_sel.astype(np.int64),
        src_offset=o_idx.astype(np.int64),
    )


def assign(
    targets: Sequence,
    levels: Sequence[tuple[int, Sequence[Sequence[float]]]],
    img_w: int,
    img_h: int,
    threshold: float = 4.0,
) -> list[LevelAssignment]:
    """`Yolov5LabelAssigner.__call__` (yv5.py:298-319): levels = [(stride, anchors_px)] for ll, ml, hl."""
    boxes, labels, offs = flatten_targets(targets)
    return [assign_level(boxes, labels, offs, s, a, img_w, img_h, threshold) for (s, a) in levels]
