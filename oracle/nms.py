"""Oracle (test infrastructure): restatement of `non_max_suppression` (kod/core/nms.py:9-75).

Per image: rows with obj > conf_thres (:33); conf_j = cls_j * obj (:40); multi-label candidates =
every (row, j) with conf_j > conf_thres in row-major `nonzero` order (:46-48) or, for a single
class, the best-class path (:49-51); optional class filter (:54-55); top-`max_nms` by confidence
when n > 30000 (:61-63); boxes + cls*4096 (:66-68); greedy NMS (torchvision.ops.nms, restated in
oracle/c/kod_oracle.c); first 300 kept (:70-71); output rows are the un-offset candidates (:73).

Besides the reference's output (list of (n_i,6) arrays) this also returns the keep indices into the
candidate list handed to NMS, which is what "NMS keep indices" means in the parity tests.

Ties in the `n > max_nms` branch: the reference uses `argsort(descending=True)` (not declared
stable); this oracle sorts stably (ties -> lower candidate index first).
"""
from __future__ import annotations

import ctypes
import os
from typing import NamedTuple, Optional, Sequence

import numpy as np

MAX_WH = 4096
MAX_DET = 300
MAX_NMS = 30000

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libkod_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        lib.kod_oracle_nms_greedy.restype = ctypes.c_int64
        lib.kod_oracle_nms_greedy.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p]
        lib.kod_oracle_bbiou.restype = None
        lib.kod_oracle_bbiou.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def greedy_nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """torchvision.ops.nms (CPU semantics) -> kept indices, score-descending, int64."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    nk = _lib().kod_oracle_nms_greedy(boxes.ctypes.data, scores.ctypes.data, n, float(iou_threshold),
                                      keep.ctypes.data)
    return keep[:nk].copy()


class ImageNMS(NamedTuple):
    dets: np.ndarray  # (n,6) float32: x1,y1,x2,y2,conf,cls  (un-offset)
    keep: np.ndarray  # (n,) int64 indices into `candidates`
    candidates: np.ndarray  # (m,6) float32 list handed to NMS (after the max_nms cut)
    cand_row: np.ndarray  # (m,) int64 source row in the (N,5+C) detection matrix
    n_before_cut: int


def candidates_for_image(x: np.ndarray, conf_thres: float, classes: Optional[Sequence[int]] = None):
    """nms.py:33-63 for one image `x` (N,5+C) float32 -> (cand (m,6), cand_row (m,), n_before_cut)."""
    x = np.asarray(x, dtype=np.float32)
    nc = x.shape[1] - 5
    # torch compares an fp32 tensor with a python float by casting the scalar to fp32
    thr32 = np.float32(conf_thres)
    rows = np.nonzero(x[:, 4] > thr32)[0]
    if rows.size == 0:
        return np.zeros((0, 6), np.float32), np.zeros((0,), np.int64), 0
    xs = x[rows]
    conf = (xs[:, 5:] * xs[:, 4:5]).astype(np.float32)
    if nc > 1:
        i, j = np.nonzero(conf > thr32)
        cand = np.concatenate([xs[i, :4], conf[i, j][:, None], j[:, None].astype(np.float32)], 1)
        cand_row = rows[i]
    else:
        j = conf.argmax(1)
        c = conf[np.arange(conf.shape[0]), j]
        m = c > thr32
        cand = np.concatenate([xs[:, :4], c[:, None], j[:, None].astype(np.float32)], 1)[m]
        cand_row = rows[m]
    if classes is not None:
        m = np.isin(cand[:, 5], np.asarray(classes, dtype=np.float32))
        cand, cand_row = cand[m], cand_row[m]
    n = cand.shape[0]
    if n > MAX_NMS:
        order = np.argsort(-cand[:, 4], kind="stable")[:MAX_NMS]
        cand, cand_row = cand[order], cand_row[order]
    return cand.astype(np.float32), cand_row.astype(np.int64), n


def nms_image(x: np.ndarray, conf_thres: float, nms_thres: float, classes=None, backend: str = "c") -> ImageNMS:
    """backend "c": the restated greedy kernel; "torchvision": the third-party op the reference itself
    calls (kod/core/nms.py:69) – used by bench.py's reference arm when torchvision is importable."""
    cand, cand_row, n0 = candidates_for_image(x, conf_thres, classes)
    if cand.shape[0] == 0:
        return ImageNMS(np.zeros((0, 6), np.float32), np.zeros((0,), np.int64), cand, cand_row, n0)
    off = (cand[:, 5:6] * np.float32(MAX_WH)).astype(np.float32)
    boxes = (cand[:, :4] + off).astype(np.float32)
    if backend == "torchvision":
        import torch
        import torchvision

        keep = torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(np.ascontiguousarray(cand[:, 4])),
                                   nms_thres).numpy()[:MAX_DET]
    else:
        keep = greedy_nms(boxes, cand[:, 4], nms_thres)[:MAX_DET]
    return ImageNMS(cand[keep], keep, cand, cand_row, n0)


def non_max_suppression(detections, conf_thres: float = 0.25, nms_thres: float = 0.45, classes=None,
                        backend: str = "c"):
    """(B,N,5+C) -> list of ImageNMS (use `.dets` for the reference's return value)."""
    det = np.asarray(detections, dtype=np.float32)
    return [nms_image(det[b], conf_thres, nms_thres, classes, backend) for b in range(det.shape[0])]
