"""CPU: host-side logic that needs no GPU – level descriptors, target packing, config structs, synthetic data."""
import numpy as np
import pytest
import torch

from object_detection_cib_b200 import _host as H
from object_detection_cib_b200 import _lib as L
from object_detection_cib_b200 import nms as N
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.types import default_anchor_info
from oracle import assign as oa


def test_level_cfgs_round_like_the_reference():
    lv = H.level_cfgs(default_anchor_info(), 416, 416)
    assert [(c.H, c.W, c.stride, c.num_anchors) for c in lv] == [(52, 52, 8, 3), (26, 26, 16, 3), (13, 13, 32, 3)]
    for c, (s, a) in zip(lv, zip(syn.STRIDES, syn.ANCHORS_PX)):
        want = oa.scaled_anchors(a, s)
        got = np.array([[c.anchor_grid[k][0], c.anchor_grid[k][1]] for k in range(3)], np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert c.fm_w == np.float32(416 / s)


def test_planar_view_has_the_head_strides():
    heads = syn.make_head(2, 64, 3, seed=0)
    box = heads[0].box
    assert box.shape == (2, 3, 8, 8, 4) and box.stride() == (3 * 4 * 64, 4 * 64, 8, 1, 64)
    assert H._planar(box) is box  # no copy for the reference layout
    contiguous = box.contiguous()  # channel-innermost: must be re-laid out
    fixed = H._planar(contiguous)
    assert fixed.stride()[3] == 1 and torch.equal(fixed, contiguous)


def test_nms_cfg_class_mask():
    cfg = N._nms_cfg(80, 0.001, 0.6, [1, 33, 79], 300, 30000)
    assert cfg.class_mask[0] == 1 << 1 and cfg.class_mask[1] == 1 << 1 and cfg.class_mask[2] == 1 << 15
    allc = N._nms_cfg(10, 0.25, 0.45, None, 300, 30000)
    assert allc.class_mask[0] == (1 << 10) - 1 and abs(allc.iou_thres - 0.45) < 1e-15


def test_zipf_targets_are_seeded_and_clamped():
    a = syn.make_targets(8, 640, 10, seed=5)
    b = syn.make_targets(8, 640, 10, seed=5)
    for (ba, la), (bb, lb) in zip(a, b):
        assert torch.equal(ba, bb) and torch.equal(la, lb)
        assert ba.numel() == 0 or (ba.min() >= 0 and ba.max() <= 639)
    labels = torch.cat([l for _, l in syn.make_targets(400, 640, 10, seed=6, n_min=10)])
    assert labels.bincount(minlength=10)[0] > labels.bincount(minlength=10)[9] * 3  # long tail


def test_struct_sizes_match_the_header():
    # kodb_level_cfg: 4 i32 + 2*8*2 f32 + 3 f32; kodb_head_level: 3 ptr + 8 i64; kodb_targets: 3 ptr + 4 i32
    import ctypes as C
    assert C.sizeof(L.LevelCfg) == 16 + 128 + 12
    assert C.sizeof(L.HeadLevel) == 24 + 64
    assert C.sizeof(L.Targets) == 24 + 16
    assert C.sizeof(L.NmsCfg) == 4 + 4 + 8 + 4 + 4 + 4 + 4 + 16 + 4 + 4  # incl. alignment padding


def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        N.nms_padded(torch.zeros(1, 10, 15))
