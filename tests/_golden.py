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
