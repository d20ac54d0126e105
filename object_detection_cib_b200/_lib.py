"""ctypes binding of libkodb.so (C ABI declared in include/kodb.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.  The
library is built in-tree by `__graft_entry__.build()` / `make -C object_detection_cib_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

NUM_LEVELS = 3
MAX_ANCHORS = 8
MAX_CLASSES = 128
NUM_IOU_THRS = 4

OK = 0
IOU_TYPES = {"iou": 0, "giou": 1, "diou": 2, "ciou": 3}
BOXES_F32, BOXES_F64 = 0, 1

_F2 = (C.c_float * 2) * MAX_ANCHORS


class LevelCfg(C.Structure):
    _fields_ = [
        ("stride", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("num_anchors", C.c_int32),
        ("anchor_px", _F2), ("anchor_grid", _F2),
        ("fm_w", C.c_float), ("fm_h", C.c_float), ("obj_ratio", C.c_float),
    ]


class HeadLevel(C.Structure):
    _fields_ = [
        ("box", C.c_void_p), ("obj", C.c_void_p), ("cls", C.c_void_p),
        ("box_sb", C.c_int64), ("box_sa", C.c_int64), ("box_sp", C.c_int64),
        ("obj_sb", C.c_int64), ("obj_sa", C.c_int64),
        ("cls_sb", C.c_int64), ("cls_sa", C.c_int64), ("cls_sp", C.c_int64),
    ]


GradLevel = HeadLevel  # same layout, writable pointers


class Targets(C.Structure):
    _fields_ = [
        ("boxes", C.c_void_p), ("labels", C.c_void_p), ("img_offsets", C.c_void_p),
        ("boxes_dtype", C.c_int32), ("num_targets", C.c_int32), ("batch", C.c_int32),
        ("max_per_image", C.c_int32),
    ]


class LossCfg(C.Structure):
    _fields_ = [
        ("lambda_classification", C.c_float), ("lambda_localization", C.c_float),
        ("lambda_objectness", C.c_float), ("iou_type", C.c_int32), ("eps", C.c_float),
        ("num_classes", C.c_int32), ("img_w", C.c_int32), ("img_h", C.c_int32),
        ("assign_threshold", C.c_float),
    ]


class AssignOut(C.Structure):
    _fields_ = [
        ("samples", C.c_void_p), ("anchors", C.c_void_p), ("grid_y", C.c_void_p), ("grid_x", C.c_void_p),
        ("labels", C.c_void_p), ("gt_boxes", C.c_void_p), ("anchor_wh", C.c_void_p),
    ]


class HeadConvLevel(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("weight", C.c_void_p), ("bias_box", C.c_void_p), ("bias_obj", C.c_void_p),
        ("bias_cls", C.c_void_p), ("out_box", C.c_void_p), ("out_obj", C.c_void_p), ("out_cls", C.c_void_p),
        ("in_channels", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
    ]


class NmsCfg(C.Structure):
    _fields_ = [
        ("conf_thres", C.c_float), ("iou_thres", C.c_double), ("max_wh", C.c_float),
        ("max_det", C.c_int32), ("max_nms", C.c_int32), ("num_classes", C.c_int32),
        ("class_mask", C.c_uint32 * (MAX_CLASSES // 32)), ("stage_mask", C.c_int32),
    ]


class NmsOut(C.Structure):
    _fields_ = [
        ("dets", C.c_void_p), ("keep_idx", C.c_void_p), ("counts", C.c_void_p),
        ("num_candidates", C.c_void_p), ("keep_key", C.c_void_p),
    ]


class MatchOut(C.Structure):
    _fields_ = [
        ("rec_score", C.c_void_p), ("rec_cls", C.c_void_p), ("rec_tp", C.c_void_p),
        ("gt_count", C.c_void_p), ("iou_thrs", C.c_double * NUM_IOU_THRS),
    ]


# symbol -> (restype, argtypes); every symbol include/kodb.h declares must be listed here
_P = C.c_void_p
SIGNATURES = {
    "kodb_version": (C.c_int, []),
    "kodb_error_string": (C.c_char_p, [C.c_int]),
    "kodb_assign": (C.c_int, [C.POINTER(LevelCfg), C.POINTER(Targets), C.c_float, C.POINTER(AssignOut),
                              C.c_int32, _P, _P]),
    "kodb_train_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "kodb_train_status": (C.c_int, [_P, C.POINTER(C.c_int32), _P]),
    "kodb_workspace_reset": (C.c_int, [_P, C.c_int64, _P]),
    "kodb_train_tail": (C.c_int, [C.POINTER(HeadLevel), C.POINTER(LevelCfg), C.POINTER(Targets),
                                  C.POINTER(LossCfg), _P, C.POINTER(C.c_float), _P, C.c_int32,
                                  C.POINTER(GradLevel), _P, _P, _P, C.c_int64, _P]),
    "kodb_iou_aligned": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_float, _P, _P, _P, _P, _P]),
    "kodb_head_forward": (C.c_int, [C.POINTER(HeadConvLevel), C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "kodb_head_detections": (C.c_int, [C.POINTER(HeadConvLevel), C.POINTER(LevelCfg), C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, _P, _P]),
    "kodb_sgd_step": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int64), C.c_int32, _P, C.c_int32, C.c_int32, _P]),
    "kodb_decode": (C.c_int, [C.POINTER(HeadLevel), C.POINTER(LevelCfg), C.c_int32, C.c_int32, _P, _P]),
    "kodb_eval_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int64, C.c_int32]),
    "kodb_nms": (C.c_int, [_P, C.c_int32, C.c_int64, C.POINTER(NmsCfg), C.POINTER(NmsOut),
                           C.POINTER(Targets), C.POINTER(MatchOut), _P, C.c_int64, _P]),
    "kodb_eval_tail": (C.c_int, [C.POINTER(HeadLevel), C.POINTER(LevelCfg), C.c_int32, C.POINTER(NmsCfg),
                                 C.POINTER(NmsOut), C.POINTER(Targets), C.POINTER(MatchOut), _P,
                                 C.c_int64, _P]),
    "kodb_map_match": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Targets),
                                 C.POINTER(MatchOut), _P]),
    "kodb_ap_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32, C.c_int32]),
    "kodb_ap_accumulate": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P,
                                     C.c_int64, _P]),
}

# (KODB_LIB: another build of the same C ABI, for A/B timing of kernel versions on one box)
LIB_PATH = os.environ.get("KODB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libkodb.so")
_lib = None


def lib():
    """Load libkodb.so (once).  Fails loudly: there is no CPU or PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built.  Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C object_detection_cib_b200/csrc`.  object_detection_cib_b200 has no fallback path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != OK:
        msg = lib().kodb_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")
