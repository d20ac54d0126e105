"""Target ingest: collate the per-image targets ONCE, in the DataLoader, into the flat
`[boxes | labels | offsets]` blob the kernels read, so that a training / validation step uploads its ground
truth with a single H2D copy of pinned memory and no per-step packing.

* `packed_collate_fn(batch)` – drop-in for `DetectionDataModule.default_collate_fn`
  (kod/lightning/data_module.py:82-87): same `(imgs, targets, image_info)` triple, but `targets` is a
  `HostPackedTargets`.
* `HostPackedTargets` – a `Sequence[DetectionTarget]` (indexing / iteration give per-image views of the blob,
  so reference code that loops over `targets` keeps working) that also carries the flat arrays.  It implements
  `pin_memory()`, which `DataLoader(pin_memory=True)` calls on custom batch objects in the main process.
  `Yolov5Loss`, `nms.eval_tail`, `nms.nms_padded` and the mAP evaluator accept it wherever they accept the
  reference's tuple of targets (`_host.pack_targets`).

dtype rule (yv5.py:92-113): if any image carries fp64 boxes the blob is fp64 and the kernels convert after the
cxcywh + 1/stride arithmetic, exactly as for the tuple form.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .types import DetectionTarget


def _round16(n: int) -> int:
    return (n + 15) // 16 * 16


class HostPackedTargets(Sequence):
    """Flat copy of the targets of one batch: `boxes (T,4)`, `labels (T,) int64`, `offsets (B+1,) int32`,
    all views of ONE uint8 blob `[boxes | labels | offsets]` with 16-byte aligned sections.

    Built on the host by the collate function; `.to(device)` moves the blob with one copy and returns the same
    kind of object, so Lightning's `transfer_batch_to_device` (which calls `.to` on anything that has it instead of
    rebuilding the sequence element by element) carries it to the GPU intact."""

    def __init__(self, blob: torch.Tensor, counts: Sequence[int], f64: bool):
        self.blob = blob
        self.counts = [int(c) for c in counts]
        self.f64 = bool(f64)
        T, B = sum(self.counts), len(self.counts)
        bsz = 8 if f64 else 4
        self.nb_boxes = _round16(T * 4 * bsz)
        self.nb_labels = _round16(T * 8)
        self.nb_offsets = _round16((B + 1) * 4)
        if blob.numel() < max(self.nb_boxes + self.nb_labels + self.nb_offsets, 16):
            raise ValueError("blob too small for the given counts")
        bdt = torch.float64 if f64 else torch.float32
        self.boxes = blob[:T * 4 * bsz].view(bdt).view(T, 4)
        self.labels = blob[self.nb_boxes:self.nb_boxes + T * 8].view(torch.long)
        self.offsets = blob[self.nb_boxes + self.nb_labels:self.nb_boxes + self.nb_labels + (B + 1) * 4].view(torch.int32)

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_targets(cls, targets: Sequence, pin: bool = False) -> "HostPackedTargets":
        """tuple of B `DetectionTarget(boxes (n_i,4) xyxy px, labels (n_i,))` (possibly empty `(0,)`-shaped arrays,
        sample_reader.py:133) -> one blob.  Lengths come from `.shape`."""
        counts = [int(t[0].shape[0]) if t[0].dim() > 0 and t[0].numel() else 0 for t in targets]
        boxes_l = [t[0].reshape(-1, 4) for t, n in zip(targets, counts) if n]
        labels_l = [t[1].reshape(-1) for t, n in zip(targets, counts) if n]
        f64 = any(b.dtype == torch.float64 for b in boxes_l)
        T, B = sum(counts), len(counts)
        bsz = 8 if f64 else 4
        nbytes = max(_round16(T * 4 * bsz) + _round16(T * 8) + _round16((B + 1) * 4), 16)
        blob = torch.zeros(nbytes, dtype=torch.uint8)
        if pin:
            blob = blob.pin_memory()
        out = cls(blob, counts, f64)
        if T:
            bdt = torch.float64 if f64 else torch.float32
            torch.cat([b.detach().to("cpu", bdt) for b in boxes_l], 0, out=out.boxes)
            torch.cat([l.detach().to("cpu", torch.long) for l in labels_l], 0, out=out.labels)
        offs = np.zeros(B + 1, np.int32)
        np.cumsum(counts, out=offs[1:])
        out.offsets.copy_(torch.from_numpy(offs))
        return out

    def pin_memory(self) -> "HostPackedTargets":
        """DataLoader(pin_memory=True) protocol for custom batch types."""
        if self.blob.is_cuda or self.blob.is_pinned():
            return self
        return HostPackedTargets(self.blob.pin_memory(), self.counts, self.f64)

    def to(self, device=None, *args, non_blocking: bool = True, **kwargs) -> "HostPackedTargets":
        """Move the blob (ONE copy) and keep the packed form.  Having `.to` makes the object a transferable leaf for
        `lightning.pytorch.utilities.move_data_to_device` / `apply_to_collection`, which would otherwise try to
        rebuild this Sequence as `type(self)([...])`.  dtype arguments are ignored (boxes keep their precision)."""
        if device is None or isinstance(device, torch.dtype):
            return self
        device = torch.device(device)
        if device == self.blob.device:
            return self
        return HostPackedTargets(self.blob.to(device, non_blocking=non_blocking), self.counts, self.f64)

    @property
    def device(self) -> torch.device:
        return self.blob.device

    # ---- Sequence[DetectionTarget] ----------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.counts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        lo = sum(self.counts[:i])  # host-side lengths: no device sync when the blob lives on the GPU
        hi = lo + self.counts[i]
        return DetectionTarget(self.boxes[lo:hi], self.labels[lo:hi])

    @property
    def num_targets(self) -> int:
        return sum(self.counts)


def packed_collate_fn(batch):
    """Drop-in for `DetectionDataModule.default_collate_fn` (kod/lightning/data_module.py:82-87)."""
    imgs, target, image_info = zip(*batch)
    return torch.stack(imgs, dim=0), HostPackedTargets.from_targets(target), image_info


# ---- box transforms of the augmentation pipeline ---------------------------------------------------
# The reference's mosaic / random-perspective / flip / mixup steps each rewrite the boxes of a sample with numpy
# (kod/data/mosaic.py:133-158, kod/data/augmentations/default.py:250-276, :318-345, :385-408) and hand the result to the
# collate function as per-image arrays.  These restate the BOX side of those steps (the pixel side stays with the
# reference's cv2 / albumentations calls) on plain arrays, so that a DataLoader worker can go from the raw sample boxes
# to the packed blob of `packed_collate_fn` without intermediate AugmentedSample objects.  CPU code by nature (it runs in
# the DataLoader workers, per sample, on a handful of boxes); float64 like the reference's numpy pipeline.
def box_candidates(orig_xyxy_t: np.ndarray, proc_xyxy_t: np.ndarray, wh_threshold: float = 2.0,
                   aspect_ratio_threshold: float = 20.0, area_thr: float = 0.1, eps: float = 1e-7) -> np.ndarray:
    """Boxes (4, n) before / after a transform -> keep mask (kod/data/mosaic.py:11-33; default.py:193-215 uses eps=1e-16)."""
    w1, h1 = orig_xyxy_t[2] - orig_xyxy_t[0], orig_xyxy_t[3] - orig_xyxy_t[1]
    w2, h2 = proc_xyxy_t[2] - proc_xyxy_t[0], proc_xyxy_t[3] - proc_xyxy_t[1]
    ar = np.maximum(w2 / (h2 + eps), h2 / (w2 + eps))
    return (w2 > wh_threshold) & (h2 > wh_threshold) & (w2 * h2 / (w1 * h1 + eps) > area_thr) & (ar < aspect_ratio_threshold)


def mosaic_placement(i: int, xc: int, yc: int, w: int, h: int, target_size: int):
    """Tile i (0 top-left, 1 top-right, 2 bottom-left, 3 bottom-right) of a (w, h) image around the mosaic centre:
    destination window in the 2S x 2S canvas, source window in the tile, and the (padw, padh) its boxes move by
    (kod/data/mosaic.py:60-131)."""
    s2 = target_size * 2
    if i == 0:
        x1a, y1a, x2a, y2a = max(xc - w, 0), max(yc - h, 0), xc, yc
        x1b, y1b, x2b, y2b = w - (x2a - x1a), h - (y2a - y1a), w, h
    elif i == 1:
        x1a, y1a, x2a, y2a = xc, max(yc - h, 0), min(xc + w, s2), yc
        x1b, y1b, x2b, y2b = 0, h - (y2a - y1a), min(w, x2a - x1a), h
    elif i == 2:
        x1a, y1a, x2a, y2a = max(xc - w, 0), yc, xc, min(s2, yc + h)
        x1b, y1b, x2b, y2b = w - (x2a - x1a), 0, w, min(y2a - y1a, h)
    elif i == 3:
        x1a, y1a, x2a, y2a = xc, yc, min(xc + w, s2), min(s2, yc + h)
        x1b, y1b, x2b, y2b = 0, 0, min(w, x2a - x1a), min(y2a - y1a, h)
    else:
        raise ValueError("a mosaic has four tiles")
    return (x1a, y1a, x2a, y2a), (x1b, y1b, x2b, y2b), (x1a - x1b, y1a - y1b)


def mosaic_targets(tiles: Sequence, tile_shapes_wh: Sequence, center_xy, target_size: int):
    """Boxes / labels of a four-tile mosaic (kod/data/mosaic.py:133-158): every tile's boxes move by its pad, the union
    is filtered by `box_candidates` against its clipping to the 2S canvas and clipped to [0, 2S - 1].
    `tiles` = four (boxes (n_i, 4) xyxy, labels (n_i,)) pairs; a tile without boxes contributes nothing (the reference
    re-appends the previous tile's boxes in that case -- mosaic.py:137-144 never resets `bboxes_xyxy` -- and then fails
    on the length mismatch; its dataset builder only feeds it images with boxes)."""
    xc, yc = center_xy
    boxes, labels = [], []
    for i, ((b, l), (w, h)) in enumerate(zip(tiles, tile_shapes_wh)):
        b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
        if b.shape[0] == 0:
            continue
        _, _, (padw, padh) = mosaic_placement(i, xc, yc, w, h, target_size)
        b = b.copy()
        b[:, [0, 2]] = b[:, [0, 2]] + padw
        b[:, [1, 3]] = b[:, [1, 3]] + padh
        boxes.append(b)
        labels.append(np.asarray(l).reshape(-1))
    if not boxes:
        return np.zeros((0, 4), np.float64), np.zeros((0,), np.int64)
    xyxy = np.concatenate(boxes, 0)
    lab = np.concatenate(labels, 0)
    keep = box_candidates(xyxy.T, np.clip(xyxy, 0, 2 * target_size).T)
    xyxy = xyxy[keep]
    np.clip(xyxy, 0, 2 * target_size - 1, out=xyxy)
    return xyxy, lab[keep]


def affine_targets(boxes: np.ndarray, labels: np.ndarray, M: np.ndarray, out_wh, perspective: bool, scale: float,
                   wh_threshold: float = 2.0, aspect_ratio_threshold: float = 20.0, area_thr: float = 0.1):
    """Boxes through the 3x3 matrix of `random_perspective` (default.py:250-276: the four corners are transformed, the
    axis-aligned hull is clipped to the output image) and the candidate filter against the scaled originals
    (default.py:336-345)."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    labels = np.asarray(labels).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return boxes, labels
    xy = np.ones((n * 4, 3))
    xy[:, :2] = boxes[:, [0, 1, 2, 3, 0, 3, 2, 1]].reshape(n * 4, 2)  # x1y1, x2y2, x1y2, x2y1
    xy = xy @ np.asarray(M, dtype=np.float64).T
    xy = (xy[:, :2] / xy[:, 2:3] if perspective else xy[:, :2]).reshape(n, 8)
    x, y = xy[:, [0, 2, 4, 6]], xy[:, [1, 3, 5, 7]]
    proc = np.concatenate((x.min(1), y.min(1), x.max(1), y.max(1))).reshape(4, n).T
    proc[:, [0, 2]] = proc[:, [0, 2]].clip(0, out_wh[0] - 1)
    proc[:, [1, 3]] = proc[:, [1, 3]].clip(0, out_wh[1] - 1)
    keep = box_candidates(boxes.T * scale, proc.T, wh_threshold, aspect_ratio_threshold, area_thr, eps=1e-16)
    return proc[keep], labels[keep]


def flip_targets(boxes: np.ndarray, image_width: int) -> np.ndarray:
    """horizontal_flip's box side (default.py:385-397)."""
    boxes = np.asarray(boxes)
    out = boxes.copy()
    if out.shape[0] > 0:
        out[:, 2] = image_width - 1 - boxes[:, 0]
        out[:, 0] = image_width - 1 - boxes[:, 2]
    return out


def mixup_targets(t1, t2):
    """mixup's box side (default.py:400-408): the two samples' boxes and labels, concatenated."""
    return np.concatenate((np.asarray(t1[0]).reshape(-1, 4), np.asarray(t2[0]).reshape(-1, 4)), 0), np.concatenate(
        (np.asarray(t1[1]).reshape(-1), np.asarray(t2[1]).reshape(-1)), 0)
