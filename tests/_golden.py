"""Helpers to read the committed golden fixtures (tests/golden/*.npz, made by make_golden.py)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEVEL_NAMES = ("ll", "ml", "hl")


def names(prefix: str) -> list[str]:
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name: str) -> dict:
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def targets_of(g: dict, prefix: str = "t") -> list[tuple[np.ndarray, np.ndarray]]:
    n = g[f"{prefix}_n"]
    boxes, labels = g[f"{prefix}_boxes"], g[f"{prefix}_labels"]
    out, o = [], 0
    for k in n:
        out.append((boxes[o:o + k], labels[o:o + k]))
        o += int(k)
    return out


def heads_of(g: dict, num_anchors: int = 3, device="cpu", grads: bool = False):
    """Planar (B,A*P,H,W) arrays -> [(box,obj,cls)] (B,A,H,W,P) permuted views, like the reference head."""
    out = []
    for ln in LEVEL_NAMES:
        lv = []
        for part in ("box", "obj", "cls"):
            raw = torch.from_numpy(g[f"{ln}_{'d' if grads else ''}{part}"]).to(device)
            B, AP, H, W = raw.shape
            lv.append(raw.view(B, num_anchors, AP // num_anchors, H, W).permute(0, 1, 3, 4, 2))
        out.append(tuple(lv))
    return out


# ---- evaltail_* fixtures: generator arguments + checksum instead of the head logits ------------------------------
def evaltail_inputs(g: dict):
    """Regenerate the heads/targets an `evaltail_*` fixture was made from (tests/golden/make_golden.py::run_evaltail)
    and check them against the stored checksum.  Returns (heads, targets, (B, S, C), (conf, iou))."""
    from object_detection_cib_b200 import synthetic as syn

    B, S, C, seed, clutter, n_gt = (int(v) for v in g["meta"])
    obj_shift, inject, conf, iou = (float(v) for v in g["fmeta"])
    heads = syn.make_head(B, S, C, seed=seed, obj_shift=obj_shift)
    targets = syn.make_targets(B, S, C, seed=seed, n_min=max(n_gt, 0), n_max=n_gt if n_gt >= 0 else 10)
    if clutter:
        syn.light_up_neighbourhoods(heads, targets, radius=clutter)
    elif inject > 0:
        syn.inject_ground_truth(heads, targets, S, fraction=inject, seed=seed)
    parts = [t.double().sum().item() for lv in heads for t in lv]
    parts += [float(lv[2].reshape(-1)[::9973][:8].double().sum()) for lv in heads]
    if not np.array_equal(np.array(parts, np.float64), g["head_sums"]):
        raise AssertionError("regenerated head logits differ from the ones the golden fixture was made from "
                             "(different torch build?)")
    return heads, targets, (B, S, C), (conf, iou)


def evaltail_split(g: dict):
    """Per-image (keep_idx, keep_key, rows) of an `evaltail_*` fixture."""
    out, o = [], 0
    for n in g["counts"]:
        n = int(n)
        out.append((g["keep_idx"][o:o + n], g["keep_key"][o:o + n], g["rows"][o:o + n]))
        o += n
    return out


def canon_ties(rows: np.ndarray, keys: np.ndarray):
    """Order runs of EQUAL scores by candidate key.  Only for images with more than max_nms candidates: there the
    reference re-orders by `argsort(descending=True)` (nms.py:61-63), which is not stable, so the relative order of
    exactly tied candidates is implementation-defined; oracle and kernels break ties by lower key."""
    order = np.lexsort((keys, -rows[:, 4].astype(np.float64)))
    return rows[order], keys[order]
