"""Host-side plumbing between the reference-shaped Python objects and the C ABI (include/kodb.h):
level descriptors, head-tensor stride checks, ground-truth packing, stream handles, workspaces.
PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must be a CUDA tensor: object_detection_cib_b200 runs its hot path on the GPU only "
            "(no CPU fallback).")


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ---- pyramid levels -------------------------------------------------------------------------------
def level_cfgs(anchor_info, img_w: int, img_h: int, obj_ratios: Sequence[float] = (4.0, 1.0, 0.4)):
    """AnchorBoxInfo x3 + image shape -> kodb_level_cfg[3], rounding exactly where the reference rounds
    (yv5.py:184-191, :227-239, :265-268; layers.py:24-27, :47-48)."""
    arr = (L.LevelCfg * L.NUM_LEVELS)()
    for i, info in enumerate(anchor_info):
        stride = int(info.stride)
        boxes = list(info.boxes_wh)
        if not 1 <= len(boxes) <= L.MAX_ANCHORS:
            raise ValueError(f"1..{L.MAX_ANCHORS} anchors per cell supported, got {len(boxes)}")
        c = arr[i]
        c.stride = stride
        c.H = int(img_h) // stride
        c.W = int(img_w) // stride
        c.num_anchors = len(boxes)
        for k, b in enumerate(boxes):
            w, h = b[0], b[1]
            c.anchor_px[k][0] = np.float32(w)
            c.anchor_px[k][1] = np.float32(h)
            c.anchor_grid[k][0] = np.float32(w * 1 / stride)
            c.anchor_grid[k][1] = np.float32(h * 1 / stride)
        c.fm_w = np.float32(img_w / stride)
        c.fm_h = np.float32(img_h / stride)
        c.obj_ratio = float(obj_ratios[i])
    return arr


# ---- head tensors ---------------------------------------------------------------------------------
def _planar(t: torch.Tensor) -> torch.Tensor:
    """Return `t` (B,A,H,W,P) with contiguous H*W planes (the reference head layout, a-0).  Tensors in
    any other layout are re-laid out on the device (one copy); values are unchanged."""
    B, A, H, W, P = t.shape
    st = t.stride()
    if t.dtype != torch.float32:
        raise TypeError(f"head tensors must be float32 (the reference trains in fp32), got {t.dtype}")
    ok = (W == 1 or st[3] == 1) and (H == 1 or st[2] == W)
    if ok:
        return t
    return t.permute(0, 1, 4, 2, 3).contiguous().permute(0, 1, 3, 4, 2)


class Heads(NamedTuple):
    structs: object  # (HeadLevel * 3)
    tensors: list  # keeps the (possibly re-laid-out) tensors alive
    batch: int
    num_classes: int
    device: torch.device


def head_structs(net_result) -> Heads:
    arr = (L.HeadLevel * L.NUM_LEVELS)()
    keep = []
    B = C_ = None
    dev = None
    for i, lv in enumerate(net_result):
        box, obj, cls = lv[0], lv[1], lv[2]
        for t in (box, obj, cls):
            require_cuda(t, "head tensor")
        box, obj, cls = _planar(box), _planar(obj), _planar(cls)
        keep += [box, obj, cls]
        if box.shape[-1] != 4 or obj.shape[-1] != 1 or box.shape[:4] != obj.shape[:4] or box.shape[:4] != cls.shape[:4]:
            raise ValueError("head tensors must be (B,A,H,W,4), (B,A,H,W,1), (B,A,H,W,C)")
        if B is None:
            B, C_, dev = box.shape[0], cls.shape[-1], box.device
        elif box.shape[0] != B or cls.shape[-1] != C_:
            raise ValueError("inconsistent batch / class count across levels")
        h = arr[i]
        h.box, h.obj, h.cls = box.data_ptr(), obj.data_ptr(), cls.data_ptr()
        h.box_sb, h.box_sa, h.box_sp = box.stride(0), box.stride(1), box.stride(4)
        h.obj_sb, h.obj_sa = obj.stride(0), obj.stride(1)
        h.cls_sb, h.cls_sa, h.cls_sp = cls.stride(0), cls.stride(1), cls.stride(4)
    return Heads(arr, keep, int(B), int(C_), dev)


def check_grid(heads: Heads, levels) -> None:
    for i in range(L.NUM_LEVELS):
        _, A, H, W, _ = heads.tensors[3 * i].shape
        if (A, H, W) != (levels[i].num_anchors, levels[i].H, levels[i].W):
            raise ValueError(
                f"level {i}: head grid (A={A},H={H},W={W}) does not match image_shape//stride "
                f"(A={levels[i].num_anchors},H={levels[i].H},W={levels[i].W})")


def alloc_planar_like(t: torch.Tensor) -> torch.Tensor:
    """Uninitialised (B,A,H,W,P) view over a fresh contiguous [B][A][P][H][W] buffer (the layout the
    1x1 head conv's backward consumes without a copy)."""
    B, A, H, W, P = t.shape
    raw = torch.empty((B, A, P, H, W), dtype=t.dtype, device=t.device)
    return raw.permute(0, 1, 3, 4, 2)


def grad_structs(grads: Sequence[torch.Tensor]):
    arr = (L.GradLevel * L.NUM_LEVELS)()
    for i in range(L.NUM_LEVELS):
        box, obj, cls = grads[3 * i], grads[3 * i + 1], grads[3 * i + 2]
        g = arr[i]
        g.box, g.obj, g.cls = box.data_ptr(), obj.data_ptr(), cls.data_ptr()
        g.box_sb, g.box_sa, g.box_sp = box.stride(0), box.stride(1), box.stride(4)
        g.obj_sb, g.obj_sa = obj.stride(0), obj.stride(1)
        g.cls_sb, g.cls_sa, g.cls_sp = cls.stride(0), cls.stride(1), cls.stride(4)
    return arr


# ---- ground truth ---------------------------------------------------------------------------------
class PackedTargets(NamedTuple):
    struct: L.Targets
    tensors: tuple  # device tensors kept alive
    num_targets: int
    batch: int
    max_per_image: int
    counts: list  # per-image target counts (host)


def pack_targets(targets: Sequence, device) -> PackedTargets:
    """tuple of B DetectionTarget(boxes (n_i,4) xyxy px, labels (n_i,)) -> flat device arrays.

    Lengths come from `.shape` (no device sync).  fp64 boxes stay fp64: the kernels convert after the
    cxcywh + 1/stride arithmetic exactly like the reference's assignment into its fp32 buffer
    (yv5.py:92-113).  Host tensors go through ONE pinned staging buffer and one H2D copy."""
    device = torch.device(device)
    from .data import HostPackedTargets  # (data.py imports nothing from this module)

    if isinstance(targets, HostPackedTargets):
        # collated in the DataLoader (data.packed_collate_fn): ONE H2D copy of the (pinned) blob, nothing to pack
        hp = targets
        T, B = hp.num_targets, len(hp)
        dblob = hp.blob.to(device, non_blocking=True)
        bsz = 8 if hp.f64 else 4
        bdt = torch.float64 if hp.f64 else torch.float32
        boxes = dblob[:T * 4 * bsz].view(bdt).view(T, 4)
        labels = dblob[hp.nb_boxes:hp.nb_boxes + T * 8].view(torch.long)
        offsets = dblob[hp.nb_boxes + hp.nb_labels:hp.nb_boxes + hp.nb_labels + (B + 1) * 4].view(torch.int32)
        s = L.Targets()
        s.boxes = boxes.data_ptr() if T else 0
        s.labels = labels.data_ptr() if T else 0
        s.img_offsets = offsets.data_ptr()
        s.boxes_dtype = L.BOXES_F64 if hp.f64 else L.BOXES_F32
        s.num_targets = T
        s.batch = B
        s.max_per_image = max(hp.counts) if hp.counts else 0
        return PackedTargets(s, (boxes, labels, offsets, dblob), T, B, s.max_per_image, list(hp.counts))
    B = len(targets)
    counts = [int(t[0].shape[0]) if t[0].dim() > 0 and t[0].numel() else 0 for t in targets]
    T = sum(counts)
    offs = np.zeros(B + 1, np.int32)
    np.cumsum(counts, out=offs[1:])
    boxes_l = [t[0].reshape(-1, 4) for t, n in zip(targets, counts) if n]
    labels_l = [t[1].reshape(-1) for t, n in zip(targets, counts) if n]
    f64 = any(b.dtype == torch.float64 for b in boxes_l)
    bdt = torch.float64 if f64 else torch.float32
    on_dev = bool(boxes_l) and all(b.is_cuda for b in boxes_l)
    if on_dev:
        boxes = torch.cat([b.to(bdt) for b in boxes_l], 0).contiguous()
        labels = torch.cat([l.to(torch.long) for l in labels_l], 0).contiguous()
        offsets = torch.from_numpy(offs).pin_memory().to(device, non_blocking=True)
    else:
        # one pinned blob: [boxes | labels | offsets], 16-byte aligned sections
        bsz = 8 if f64 else 4
        nb_boxes = (T * 4 * bsz + 15) // 16 * 16
        nb_labels = (T * 8 + 15) // 16 * 16
        nb_offs = ((B + 1) * 4 + 15) // 16 * 16
        blob = torch.empty(max(nb_boxes + nb_labels + nb_offs, 16), dtype=torch.uint8).pin_memory()
        hb = blob[:T * 4 * bsz].view(bdt).view(T, 4)
        hl = blob[nb_boxes:nb_boxes + T * 8].view(torch.long)
        ho = blob[nb_boxes + nb_labels:nb_boxes + nb_labels + (B + 1) * 4].view(torch.int32)
        if T:
            torch.cat([b.to(bdt) for b in boxes_l], 0, out=hb)
            torch.cat([l.to(torch.long) for l in labels_l], 0, out=hl)
        ho.copy_(torch.from_numpy(offs))
        dblob = blob.to(device, non_blocking=True)
        boxes = dblob[:T * 4 * bsz].view(bdt).view(T, 4)
        labels = dblob[nb_boxes:nb_boxes + T * 8].view(torch.long)
        offsets = dblob[nb_boxes + nb_labels:nb_boxes + nb_labels + (B + 1) * 4].view(torch.int32)
    s = L.Targets()
    s.boxes = boxes.data_ptr() if T else 0
    s.labels = labels.data_ptr() if T else 0
    s.img_offsets = offsets.data_ptr()
    s.boxes_dtype = L.BOXES_F64 if f64 else L.BOXES_F32
    s.num_targets = T
    s.batch = B
    s.max_per_image = max(counts) if counts else 0
    return PackedTargets(s, (boxes, labels, offsets), T, B, s.max_per_image, counts)


def check_labels_host(targets, num_classes: int) -> None:
    """Reference behaviour for a class id outside [0, C): `target_labels_one_hot[range(n), labels] = 1`
    (loss.py:145-152) raises IndexError.  Checked here whenever the labels are still on the host (tuple of CPU
    tensors, or a collated blob that has not been moved yet) -- a handful of integers, no device sync.  Labels that
    arrive on the device are checked by the train-tail kernel instead (NaN classification loss + error flag)."""
    from .data import HostPackedTargets

    if isinstance(targets, PackedTargets):
        return
    if isinstance(targets, HostPackedTargets):
        labs = [targets.labels] if not targets.labels.is_cuda else []
    else:
        labs = [t[1].reshape(-1) for t in targets if not t[1].is_cuda and t[1].numel()]
    for l in labs:
        if l.numel() and (int(l.min()) < 0 or int(l.max()) >= num_classes):
            bad = int(l.max()) if int(l.max()) >= num_classes else int(l.min())
            raise IndexError(f"target label {bad} is out of range for {num_classes} classes")


# ---- workspaces -----------------------------------------------------------------------------------
_WORKSPACES: dict = {}
_MAX_CACHED = 8


def workspace(kind: str, device, nbytes: int, tag=None) -> torch.Tensor:
    """Zero-initialised, cached per (kind, device, stream, tag); grown on demand.  Kernels leave workspaces
    reusable FOR THE SAME LAYOUT, so the memset happens only at (re)allocation.  `tag` names the layout of the
    zero-on-entry regions (the eval workspace places them by batch size, include/kodb.h): a different tag gets a
    different buffer instead of a stale one.  The cache keeps the most recently used few buffers."""
    key = (kind, torch.device(device).index, torch.cuda.current_stream(device).cuda_stream, tag)
    ws = _WORKSPACES.pop(key, None)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
    _WORKSPACES[key] = ws  # (re)insert as most recent
    while len(_WORKSPACES) > _MAX_CACHED:
        _WORKSPACES.pop(next(iter(_WORKSPACES)))
    return ws
