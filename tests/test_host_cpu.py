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


def test_packed_collate_matches_the_reference_collate():
    """data.packed_collate_fn returns the reference's triple; the targets object is a
    Sequence[DetectionTarget] whose items equal the inputs, for fp32, fp64, empty and (0,)-shaped targets."""
    from object_detection_cib_b200.data import HostPackedTargets, packed_collate_fn
    from object_detection_cib_b200.types import DetectionTarget
    g = torch.Generator().manual_seed(3)
    targets = []
    for n in (3, 0, 7, 1):
        b = torch.rand(n, 4, generator=g) * 600
        targets.append(DetectionTarget(torch.cat([b[:, :2], b[:, :2] + 10 + b[:, 2:] / 6], 1), torch.randint(0, 10, (n,), generator=g)))
    targets[1] = DetectionTarget(torch.zeros(0), torch.zeros(0, dtype=torch.long))  # sample_reader.py:133 shape
    batch = [(torch.full((3, 8, 8), float(i)), t, ("p%d" % i, (8, 8))) for i, t in enumerate(targets)]
    imgs, packed, info = packed_collate_fn(batch)
    assert imgs.shape == (4, 3, 8, 8) and info[2][0] == "p2"
    assert isinstance(packed, HostPackedTargets) and len(packed) == 4 and packed.num_targets == 11
    assert packed.offsets.tolist() == [0, 3, 3, 10, 11] and not packed.f64
    for got, want in zip(packed, targets):
        assert torch.equal(got.boxes, want.boxes.reshape(-1, 4)) and torch.equal(got.labels, want.labels.reshape(-1))
    assert torch.equal(packed[-1].labels, targets[3].labels)
    # fp64 boxes of ONE image make the blob fp64 (the kernels convert after the cxcywh + 1/stride arithmetic)
    t64 = list(targets)
    t64[2] = DetectionTarget(targets[2].boxes.double() + 1e-9, targets[2].labels)
    p64 = HostPackedTargets.from_targets(t64)
    assert p64.f64 and p64.boxes.dtype == torch.float64 and torch.equal(p64[2].boxes, t64[2].boxes)
    assert torch.equal(p64[0].boxes, targets[0].boxes.double())
    # an all-empty batch
    pe = HostPackedTargets.from_targets([targets[1], targets[1]])
    assert len(pe) == 2 and pe.num_targets == 0 and pe.offsets.tolist() == [0, 0, 0]


def test_packed_targets_work_with_a_dataloader():
    from object_detection_cib_b200.data import HostPackedTargets, packed_collate_fn
    from object_detection_cib_b200.types import DetectionTarget

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 6

        def __getitem__(self, i):
            n = i % 3
            return torch.zeros(3, 4, 4), DetectionTarget(torch.full((n, 4), float(i)), torch.full((n,), i, dtype=torch.long)), None

    dl = torch.utils.data.DataLoader(DS(), batch_size=3, collate_fn=packed_collate_fn, num_workers=0)
    batches = list(dl)
    assert len(batches) == 2
    _, t, _ = batches[1]
    assert isinstance(t, HostPackedTargets) and t.counts == [0, 1, 2] and float(t[2].boxes[0, 0]) == 5.0


# ---- Lightning's batch transfer: apply_to_collection must not rebuild the packed targets ---------------
def _apply_to_collection(data, dtype, function, *args, **kwargs):
    """lightning_utilities.core.apply_func.apply_to_collection, restated for the container types a detection batch
    holds (lightning is not installed here): a leaf of `dtype` goes through `function`; mappings, namedtuples and
    generic Sequences are REBUILT as `type(data)(...)` from their transformed elements."""
    from collections.abc import Mapping, Sequence
    if isinstance(data, dtype):
        return function(data, *args, **kwargs)
    elem_type = type(data)
    if isinstance(data, Mapping):
        return elem_type({k: _apply_to_collection(v, dtype, function, *args, **kwargs) for k, v in data.items()})
    is_nt = isinstance(data, tuple) and hasattr(data, "_fields")
    if is_nt or (isinstance(data, Sequence) and not isinstance(data, str)):
        out = [_apply_to_collection(d, dtype, function, *args, **kwargs) for d in data]
        return elem_type(*out) if is_nt else elem_type(out)
    return data


def test_packed_targets_survive_lightning_batch_transfer():
    from abc import ABC
    from object_detection_cib_b200.data import HostPackedTargets, packed_collate_fn
    from object_detection_cib_b200.types import DetectionTarget

    class Transferable(ABC):
        @classmethod
        def __subclasshook__(cls, subclass):
            return callable(getattr(subclass, "to", None)) if cls is Transferable else NotImplemented

    dev = torch.device("cpu")
    moved = []

    def batch_to(data):  # lightning.fabric.utilities.apply_func.move_data_to_device::batch_to
        kwargs = {"non_blocking": True} if isinstance(data, torch.Tensor) else {}
        out = data.to(dev, **kwargs)
        moved.append(type(data).__name__)
        return out

    targets = [DetectionTarget(torch.rand(n, 4), torch.randint(0, 10, (n,))) for n in (2, 0, 3)]
    batch = packed_collate_fn([(torch.zeros(3, 4, 4), t, ("p", (4, 4))) for t in targets])
    out = _apply_to_collection(batch, Transferable, batch_to)
    imgs, packed, info = out
    assert isinstance(packed, HostPackedTargets) and packed.counts == [2, 0, 3]  # a leaf: moved with ONE .to()
    assert moved.count("HostPackedTargets") == 1 and info == batch[2]
    for got, want in zip(packed, targets):
        assert torch.equal(got.boxes, want.boxes) and torch.equal(got.labels, want.labels)
    # .to() keeps the packed form on another device type too (meta: no data, just the plumbing)
    m = packed.to("meta")
    assert isinstance(m, HostPackedTargets) and m.device.type == "meta" and m[2].boxes.shape == (3, 4)
    assert packed.to(torch.float32) is packed and packed.to("cpu") is packed
    # without `.to` the rebuild path would have raised: HostPackedTargets([...]) has no such constructor
    with pytest.raises(TypeError):
        HostPackedTargets([t for t in packed])


def test_out_of_range_labels_raise_like_the_reference():
    """loss.py:145-152 indexes a one-hot matrix with the labels: an id outside [0, C) is an IndexError there."""
    from object_detection_cib_b200 import _host as H
    t = [(torch.tensor([[1.0, 2, 30, 40]]), torch.tensor([10])), (torch.zeros(0, 4), torch.zeros(0, dtype=torch.long))]
    with pytest.raises(IndexError, match="label"):
        H.check_labels_host(t, 10)
    with pytest.raises(IndexError, match="label"):
        H.check_labels_host([(torch.tensor([[1.0, 2, 30, 40]]), torch.tensor([-1]))], 10)
    H.check_labels_host([(torch.tensor([[1.0, 2, 30, 40]]), torch.tensor([9]))], 10)


# ---- box transforms of the augmentation pipeline against the real reference ---------------------------
def test_augment_box_transforms_match_the_reference():
    """mosaic (mosaic.py:133-158), random perspective (default.py:250-276, :336-345), horizontal flip (:385-397) and mixup
    (:400-408): boxes and labels bit for bit against the reference functions (tests/golden/augment_boxes.npz), and the
    result collates into the packed blob."""
    import _golden as G
    from object_detection_cib_b200 import data as D
    from object_detection_cib_b200.types import DetectionTarget
    g = G.load("augment_boxes")
    collated = []
    for case in range(4):
        tiles = [(g[f"mosaic{case}/tile{i}_boxes"], g[f"mosaic{case}/tile{i}_labels"]) for i in range(4)]
        shapes = [tuple(int(v) for v in g[f"mosaic{case}/tile{i}_wh"]) for i in range(4)]
        xc, yc = (int(v) for v in g[f"mosaic{case}/center"])
        mb, ml = D.mosaic_targets(tiles, shapes, (xc, yc), 320)
        assert np.array_equal(mb, g[f"mosaic{case}/out_boxes"]) and np.array_equal(ml, g[f"mosaic{case}/out_labels"])
        persp, scale = g[f"affine{case}/perspective_scale"]
        ab, al = D.affine_targets(mb, ml, g[f"affine{case}/M"], tuple(int(v) for v in g[f"affine{case}/out_wh"]),
                                  bool(persp), float(scale))
        assert np.array_equal(ab, g[f"affine{case}/out_boxes"]) and np.array_equal(al, g[f"affine{case}/out_labels"])
        fb = D.flip_targets(ab, int(g[f"affine{case}/out_wh"][0]))
        assert np.array_equal(fb, g[f"flip{case}/out_boxes"])
        if case:
            xb, xl = D.mixup_targets((g[f"flip{case - 1}/out_boxes"], g[f"affine{case - 1}/out_labels"]), (fb, al))
            assert np.array_equal(xb, g[f"mixup{case}/out_boxes"]) and np.array_equal(xl, g[f"mixup{case}/out_labels"])
            fb, al = xb, xl
        collated.append((torch.zeros(3, 4, 4), DetectionTarget(torch.from_numpy(fb), torch.from_numpy(al.astype(np.int64))), None))
    _, packed, _ = D.packed_collate_fn(collated)
    assert packed.f64 and packed.num_targets == sum(int(c[1].labels.shape[0]) for c in collated)
    # a tile without boxes contributes nothing; the placement windows tile the canvas around the centre
    b, l = D.mosaic_targets([(np.zeros((0, 4)), np.zeros((0,)))] * 4, [(100, 100)] * 4, (320, 320), 320)
    assert b.shape == (0, 4) and l.shape == (0,)
    (x1a, y1a, x2a, y2a), (x1b, y1b, x2b, y2b), (padw, padh) = D.mosaic_placement(0, 300, 280, 200, 150, 320)
    assert (x2a - x1a, y2a - y1a) == (x2b - x1b, y2b - y1b) and (padw, padh) == (x1a - x1b, y1a - y1b)
