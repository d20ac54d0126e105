"""Seeded synthetic inputs for the detection tail.

Head logits are produced in the *physical* layout the reference's head convs emit,
`(B, A*P, H, W)` contiguous, and exposed as the `(B, A, H, W, P)` permuted view that
`Rearrange("b (a p) h w -> b a h w p")` yields (kod/nn/heads/yolov5.py:29-32, :77-80, :125-128).
Logit statistics follow the head's own bias init (kod/nn/heads/yolov5.py:65-66, :113-114) so the eval
tail sees the dense-candidate regime; targets follow the dataset builder's Zipf(1.01) class
distribution (kod/data/builder.py:235-237).
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence

import torch

STRIDES = (8, 16, 32)
# kod/configs/anchor_boxes/voc_s{8,16,32}.yaml == kod/test_utils/anchor_boxes.py:6-31
ANCHORS_PX = (
    ((10, 13), (16, 30), (33, 23)),
    ((30, 61), (62, 45), (59, 119)),
    ((116, 90), (156, 198), (373, 326)),
)
# README.md:66-75 train-set instance counts -> class loss weights sum(n)/n_c (trainer.py:54-60)
README_INSTANCE_COUNTS = (12982, 6918, 2409, 2663, 1837, 1829, 1096, 1009, 941, 883)


class HeadLevel(NamedTuple):
    box: torch.Tensor  # (B,A,H,W,4) view
    obj: torch.Tensor  # (B,A,H,W,1) view
    cls: torch.Tensor  # (B,A,H,W,C) view


def planar_view(raw: torch.Tensor, num_anchors: int) -> torch.Tensor:
    """(B, A*P, H, W) contiguous -> (B, A, H, W, P) view (no copy)."""
    B, AP, H, W = raw.shape
    P = AP // num_anchors
    return raw.view(B, num_anchors, P, H, W).permute(0, 1, 3, 4, 2)


def zipf_pmf(num_classes: int, a: float = 1.01) -> torch.Tensor:
    k = torch.arange(1, num_classes + 1, dtype=torch.float64)
    p = k.pow(-a)
    return p / p.sum()


def make_head(batch: int, img_size: int, num_classes: int, seed: int = 2023, device="cpu",
              num_anchors: int = 3, strides: Sequence[int] = STRIDES, obj_shift: float = 0.0,
              img_h: Optional[int] = None) -> list[HeadLevel]:
    """Random head logits for ll/ml/hl.  `obj_shift` moves the objectness mean (candidate density)."""
    img_w = img_size
    img_h = img_size if img_h is None else img_h
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    levels = []
    for s in strides:
        H, W = img_h // s, img_w // s
        b_obj = math.log(8 / (640 / s) ** 2) + obj_shift
        b_cls = math.log(0.6 / (num_classes - 0.99999))
        box = torch.randn(batch, num_anchors * 4, H, W, generator=gen, device=dev)
        obj = torch.randn(batch, num_anchors * 1, H, W, generator=gen, device=dev) * 1.5 + b_obj
        cls = torch.randn(batch, num_anchors * num_classes, H, W, generator=gen, device=dev) + b_cls
        levels.append(HeadLevel(planar_view(box, num_anchors), planar_view(obj, num_anchors),
                                planar_view(cls, num_anchors)))
    return levels


def make_targets(batch: int, img_size: int, num_classes: int, seed: int = 2023, n_min: int = 0,
                 n_max: int = 10, size_lo: float = 4.0, size_hi: float = 204.0,
                 dtype=torch.float32, empty_images: Sequence[int] = ()) -> list[tuple[torch.Tensor, torch.Tensor]]:
    """Per image n~U{n_min..n_max} boxes: centres U(0,S)^2, sizes U(lo,hi)^2, xyxy clamped to
    [0,S-1]; labels ~ Zipf(1.01).  Always generated on the CPU (host-side ingest format)."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    pmf = zipf_pmf(num_classes)
    out = []
    for b in range(batch):
        n = int(torch.randint(n_min, n_max + 1, (1,), generator=gen))
        if b in empty_images:
            n = 0
        c = torch.rand(n, 2, generator=gen, dtype=torch.float64) * img_size
        wh = size_lo + torch.rand(n, 2, generator=gen, dtype=torch.float64) * (size_hi - size_lo)
        xyxy = torch.cat((c - wh / 2, c + wh / 2), 1).clamp(0, img_size - 1)
        labels = torch.multinomial(pmf, n, replacement=True, generator=gen) if n else torch.zeros(0, dtype=torch.long)
        out.append((xyxy.to(dtype), labels.to(torch.long)))
    return out


def inject_ground_truth(heads: Sequence[HeadLevel], targets, img_size: int, fraction: float = 0.3,
                        seed: int = 7, strides: Sequence[int] = STRIDES, anchors_px=ANCHORS_PX) -> None:
    """Make a fraction of the GT boxes detectable: write logits that decode to (nearly)
    the GT box with high obj/cls confidence at the cell holding the box centre, on the level/anchor
    whose shape fits best.  In-place on the head tensors (eval-side inputs only)."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    for b, (boxes, labels) in enumerate(targets):
        for n in range(boxes.shape[0]):
            if float(torch.rand(1, generator=gen)) >= fraction:
                continue
            x1, y1, x2, y2 = (float(v) for v in boxes[n])
            w, h = max(x2 - x1, 1.0), max(y2 - y1, 1.0)
            cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
            best = None
            for li, s in enumerate(strides):
                for ai, (aw, ah) in enumerate(anchors_px[li]):
                    r = max(w / aw, aw / w, h / ah, ah / h)
                    if best is None or r < best[0]:
                        best = (r, li, ai)
            r, li, ai = best
            if r >= 3.9:
                continue
            s = strides[li]
            aw, ah = anchors_px[li][ai]
            H, W = heads[li].obj.shape[2], heads[li].obj.shape[3]
            gx, gy = min(int(cx / s), W - 1), min(int(cy / s), H - 1)
            # invert xy = (2*sig + g - 0.5)*s and wh = (2*sig)^2*anchor
            sx = min(max((cx / s - gx + 0.5) / 2, 1e-3), 1 - 1e-3)
            sy = min(max((cy / s - gy + 0.5) / 2, 1e-3), 1 - 1e-3)
            sw = min(max(math.sqrt(w / aw) / 2, 1e-3), 1 - 1e-3)
            sh = min(max(math.sqrt(h / ah) / 2, 1e-3), 1 - 1e-3)
            logit = lambda p: math.log(p / (1 - p))
            heads[li].box[b, ai, gy, gx, :] = torch.tensor([logit(sx), logit(sy), logit(sw), logit(sh)],
                                                          device=heads[li].box.device)
            heads[li].obj[b, ai, gy, gx, 0] = 4.0
            heads[li].cls[b, ai, gy, gx, int(labels[n])] = 4.0


def light_up_neighbourhoods(heads: Sequence[HeadLevel], targets, radius: int = 2, seed: int = 1,
                            strides: Sequence[int] = STRIDES, anchors_px=ANCHORS_PX) -> None:
    """Detector-like clutter: every ground-truth box lights up the (2*radius+1)^2 cells around its centre on every
    anchor that fits it, each decoding to (nearly) the box with high objectness / class confidence.  NMS then has to
    scan thousands of heavily overlapping candidates for its 300 survivors (several extraction batches, score floors
    that run dry).  In-place on the head tensors."""
    gen = torch.Generator()
    gen.manual_seed(seed)
    lg = lambda q: math.log(q / (1 - q))
    for b, (boxes, labels) in enumerate(targets):
        for n in range(boxes.shape[0]):
            x1, y1, x2, y2 = (float(v) for v in boxes[n])
            cx, cy, w, h = (x1 + x2) / 2, (y1 + y2) / 2, max(x2 - x1, 1.0), max(y2 - y1, 1.0)
            for li, s in enumerate(strides):
                H, W = heads[li].obj.shape[2], heads[li].obj.shape[3]
                gx, gy = int(cx / s), int(cy / s)
                for ai, (aw, ah) in enumerate(anchors_px[li]):
                    if max(w / aw, aw / w, h / ah, ah / h) >= 4:
                        continue
                    sw = min(max(math.sqrt(w / aw) / 2, 1e-3), 1 - 1e-3)
                    sh = min(max(math.sqrt(h / ah) / 2, 1e-3), 1 - 1e-3)
                    for dy in range(-radius, radius + 1):
                        for dx in range(-radius, radius + 1):
                            yy, xx = gy + dy, gx + dx
                            if not (0 <= yy < H and 0 <= xx < W):
                                continue
                            sx = min(max((cx / s - xx + 0.5) / 2, 0.02), 0.98)
                            sy = min(max((cy / s - yy + 0.5) / 2, 0.02), 0.98)
                            dev = heads[li].box.device
                            heads[li].box[b, ai, yy, xx, :] = torch.tensor([lg(sx), lg(sy), lg(sw), lg(sh)], device=dev)
                            heads[li].obj[b, ai, yy, xx, 0] = 3.0 - 0.5 * (abs(dx) + abs(dy)) + 0.3 * float(torch.randn(1, generator=gen))
                            heads[li].cls[b, ai, yy, xx, int(labels[n])] = 3.0 + 0.3 * float(torch.randn(1, generator=gen))


def readme_class_weights() -> list[float]:
    """`sum(n)/n_c` in fp32 (kod/lightning/tasks/trainer.py:54-60) for the README instance counts."""
    w = torch.tensor(README_INSTANCE_COUNTS, dtype=torch.float32)
    return (w.sum() / w).tolist()
