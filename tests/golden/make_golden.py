"""Generate the golden fixtures by running the REAL reference (`/root/reference`, imported with the
stubs of `_ref_harness.py`) on small seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Writes tests/golden/{assign,loss,decode,nms,evaltail}_*.npz.  Each file stores the inputs and the reference's
outputs, so the GPU box (which has no /root/reference) can check both the oracle and the CUDA path.
(`evaltail_*` fixtures store the generator arguments + a checksum of the head logits instead of the logits.)
No fixture exists for the mAP stage: its arithmetic lives in pycocotools, which is absent
("parity unpinned", see oracle/coco_map.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _ref_harness  # noqa: E402

assert _ref_harness.install(), "reference tree not found"

import torchvision  # noqa: E402
from kod.core.anchors.info import AnchorBoxInfo  # noqa: E402
from kod.core.bbox.iou import IoUCalculator, IoUType  # noqa: E402
from kod.core.label_assignment.yv5 import AssignmentAnchorInfo, Yolov5LabelAssigner  # noqa: E402
from kod.core.types import FeatureShape  # noqa: E402
from kod.data.detection import DetectionTarget  # noqa: E402
from kod.lightning.experiments.yv5_baseline.exp import DefaultYolov5Experiment  # noqa: E402
from kod.lightning.experiments.yv5_baseline.loss import Yolov5Loss, Yolov5LossParams  # noqa: E402
from kod.lightning.experiments.yv5_baseline.type_defs import LayerwiseAnchorInfo  # noqa: E402
from kod.nn.heads.types import DetectionHeadResult  # noqa: E402
from kod.nn.networks.yolov5 import Yolov5NetworkResult  # noqa: E402
import kod.core.nms as ref_nms_mod  # noqa: E402

from object_detection_cib_b200 import synthetic as syn  # noqa: E402

LEVEL_NAMES = ("ll", "ml", "hl")


def anchor_infos():
    return [AnchorBoxInfo(stride=s, boxes_wh=[FeatureShape(width=w, height=h) for (w, h) in a])
            for s, a in zip(syn.STRIDES, syn.ANCHORS_PX)]


def ref_targets(targets):
    return tuple(DetectionTarget(boxes=b, labels=l) for b, l in targets)


def pack_targets(targets, prefix="t"):
    d = {f"{prefix}_n": np.array([t[0].shape[0] for t in targets], np.int64)}
    boxes = [t[0].reshape(-1, 4) for t in targets]
    labels = [t[1].reshape(-1) for t in targets]
    dtype = boxes[0].dtype if boxes else torch.float32
    d[f"{prefix}_boxes"] = torch.cat(boxes, 0).numpy() if boxes else np.zeros((0, 4), np.float32)
    d[f"{prefix}_labels"] = torch.cat(labels, 0).numpy().astype(np.int64) if labels else np.zeros((0,), np.int64)
    return d


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


# ------------------------------------------------------------------------------------------------
# target assignment
# ------------------------------------------------------------------------------------------------
def run_assign(name, targets, img_w, img_h, threshold=4.0):
    infos = anchor_infos()
    assigner = Yolov5LabelAssigner(AssignmentAnchorInfo(*infos), threshold=threshold)
    res = assigner(FeatureShape(width=img_w, height=img_h), ref_targets(targets))
    out = pack_targets(targets)
    out["img_wh"] = np.array([img_w, img_h], np.int64)
    out["threshold"] = np.array([threshold], np.float64)
    for ln, r in zip(LEVEL_NAMES, res):
        out[f"{ln}_samples"] = r.indices.samples.numpy()
        out[f"{ln}_anchors"] = r.indices.anchors.numpy()
        out[f"{ln}_grid_y"] = r.indices.grid_y.numpy()
        out[f"{ln}_grid_x"] = r.indices.grid_x.numpy()
        out[f"{ln}_labels"] = r.labels.numpy()
        out[f"{ln}_gt_boxes"] = r.gt_boxes.numpy().reshape(-1, 4)
        out[f"{ln}_anchor_wh"] = r.anchors.numpy().reshape(-1, 2)
    save(name, **out)


def assign_cases():
    # a1: the SURVEY A.1-7 example (two identical GT in one cell, one GT in a second image)
    t = [(torch.tensor([[100.0, 100, 160, 170], [100, 100, 160, 170]]), torch.tensor([1, 3])),
         (torch.tensor([[300.0, 300, 420, 380]]), torch.tensor([0]))]
    run_assign("assign_kat", t, 640, 640)
    # a2: random cfg-1 shaped batch with one empty image
    run_assign("assign_rand640", syn.make_targets(8, 640, 10, seed=2024, empty_images=(3,)), 640, 640)
    # a3: the reference's default 416 input (52/26/13 grids), fp64 boxes as produced by the numpy pipeline
    run_assign("assign_rand416_f64", syn.make_targets(5, 416, 10, seed=2025, n_max=20, dtype=torch.float64,
                                                      size_hi=300.0), 416, 416)
    # a4: hand-made edge cases: centres on cell boundaries (duplicate rows), centres within one cell of the
    # border (neighbour masks off), zero-size / huge boxes (filtered by every anchor), non-square image
    e = [
        (torch.tensor([[64.0, 64, 96, 96],      # centre (80,80): on the ll/ml boundary grid
                       [0.0, 0, 14, 14],        # centre (7,7): < 1 cell from the border on every level
                       [300.0, 200, 300, 200],  # zero size
                       [0.0, 0, 639, 479],      # huge
                       [630.0, 470, 639, 479],  # bottom-right corner
                       [100.5, 100.25, 130.5, 121.75]]), torch.tensor([0, 9, 2, 3, 4, 5])),
        (torch.zeros(0, 4), torch.zeros(0, dtype=torch.long)),
        (torch.tensor([[320.0, 240, 352, 272], [320.0, 240, 352, 272], [322.0, 241, 350, 271]]),
         torch.tensor([7, 7, 1])),
    ]
    run_assign("assign_edges", e, 640, 480)
    # a5: every image empty
    z = [(torch.zeros(0, 4), torch.zeros(0, dtype=torch.long)) for _ in range(3)]
    run_assign("assign_empty", z, 640, 640)
    # a6: a different threshold
    run_assign("assign_thr2", syn.make_targets(4, 320, 10, seed=2026, n_max=12), 320, 320, threshold=2.0)


# ------------------------------------------------------------------------------------------------
# loss forward + backward
# ------------------------------------------------------------------------------------------------
def net_result(heads, requires_grad=False):
    lv = []
    for h in heads:
        ts = [t.detach().clone() for t in h]  # clones of the permuted views keep the planar strides
        if requires_grad:
            for t in ts:
                t.requires_grad_(True)
        lv.append(DetectionHeadResult(*ts))
    return Yolov5NetworkResult(*lv)


def run_loss(name, B, S, C, targets, weights, iou_type="ciou", seed=11, hp=None):
    heads = syn.make_head(B, S, C, seed=seed, obj_shift=3.0)
    infos = anchor_infos()
    assigner = Yolov5LabelAssigner(AssignmentAnchorInfo(*infos))
    hp = hp or Yolov5LossParams.get_default()
    loss_mod = Yolov5Loss(assigner, hp, IoUCalculator(IoUType(iou_type), eps=1e-7), weights=weights)
    nr = net_result(heads, requires_grad=True)
    res = loss_mod(FeatureShape(width=S, height=S), nr, ref_targets(targets))
    total = B * (res.localization + res.classification + res.objectness)  # exp.py:124-130
    total.backward()
    out = pack_targets(targets)
    out["meta"] = np.array([B, S, C], np.int64)
    out["iou_type"] = np.array([iou_type])
    out["hparams"] = np.array(list(hp), np.float64)
    out["weights"] = np.array(weights if weights is not None else [], np.float32)
    out["losses"] = np.array([res.localization.item(), res.objectness.item(), res.classification.item()],
                             np.float32)
    for ln, h, r in zip(LEVEL_NAMES, heads, nr):
        for part, src, leaf in zip(("box", "obj", "cls"), h, r):
            # store in the physical planar layout (B, A*P, H, W)
            out[f"{ln}_{part}"] = src.permute(0, 1, 4, 2, 3).contiguous().numpy().reshape(
                B, -1, src.shape[2], src.shape[3])
            g = leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)
            out[f"{ln}_d{part}"] = g.permute(0, 1, 4, 2, 3).contiguous().numpy().reshape(
                B, -1, src.shape[2], src.shape[3])
    save(name, **out)


def loss_cases():
    w10 = syn.readme_class_weights()
    t = syn.make_targets(3, 96, 10, seed=31, n_min=1, n_max=6, size_lo=4, size_hi=90)
    # force a duplicate cell: repeat the first GT of image 0 with another label
    b0, l0 = t[0]
    t[0] = (torch.cat((b0, b0[:1]), 0), torch.cat((l0, torch.tensor([(int(l0[0]) + 1) % 10])), 0))
    run_loss("loss_ciou_w", 3, 96, 10, t, w10)
    run_loss("loss_ciou_now", 2, 128, 4, syn.make_targets(2, 128, 4, seed=32, n_min=2, n_max=5, size_hi=120),
             None, seed=12)
    for it in ("iou", "giou", "diou"):
        run_loss(f"loss_{it}", 2, 64, 3, syn.make_targets(2, 64, 3, seed=33, n_min=2, n_max=4, size_hi=60),
                 [1.0, 2.0, 3.0], iou_type=it, seed=13)
    # empty level(s): only one tiny box -> ml/hl unassigned -> NaN loc/cls
    tiny = [(torch.tensor([[10.0, 10, 22, 24]]), torch.tensor([2])), (torch.zeros(0, 4), torch.zeros(0, dtype=torch.long))]
    run_loss("loss_empty_levels", 2, 96, 10, tiny, w10, seed=14)
    # non-multiple-of-4 grids (13x13 / 6x6 / 3x3) and non-default lambdas
    hp = Yolov5LossParams(0.3, 0.07, 0.9, 3.0, 1.5, 0.5)
    run_loss("loss_s104", 2, 104, 10, syn.make_targets(2, 104, 10, seed=34, n_min=2, n_max=5, size_hi=100),
             w10, seed=15, hp=hp)


# ------------------------------------------------------------------------------------------------
# decode
# ------------------------------------------------------------------------------------------------
class _Exp:
    """Only what `get_detections` touches (exp.py:70-102)."""

    def __init__(self):
        self.anchor_info = LayerwiseAnchorInfo(*anchor_infos())

    get_detections = DefaultYolov5Experiment.get_detections


def run_decode(name, B, S, C, seed, obj_shift=0.0):
    heads = syn.make_head(B, S, C, seed=seed, obj_shift=obj_shift)
    det = _Exp().get_detections(FeatureShape(width=S, height=S), net_result(heads))
    out = {"meta": np.array([B, S, C], np.int64), "detections": det.numpy()}
    for ln, h in zip(LEVEL_NAMES, heads):
        for part, src in zip(("box", "obj", "cls"), h):
            out[f"{ln}_{part}"] = src.permute(0, 1, 4, 2, 3).contiguous().numpy().reshape(
                B, -1, src.shape[2], src.shape[3])
    save(name, **out)
    return det


def decode_cases():
    run_decode("decode_s96", 2, 96, 10, seed=41)
    run_decode("decode_s104", 2, 104, 3, seed=42)  # 13x13 / 6x6 / 3x3 planes


# ------------------------------------------------------------------------------------------------
# NMS
# ------------------------------------------------------------------------------------------------
def run_nms(name, det, conf, iou, classes=None):
    rec = []
    real = torchvision.ops.nms

    def spy(boxes, scores, thr):
        keep = real(boxes, scores, thr)
        rec.append((boxes.clone(), scores.clone(), keep.clone()))
        return keep

    torchvision.ops.nms = spy
    try:
        res = ref_nms_mod.non_max_suppression(det.clone(), conf_thres=conf, nms_thres=iou, classes=classes)
    finally:
        torchvision.ops.nms = real
    out = {"detections": det.numpy(), "conf_iou": np.array([conf, iou], np.float64),
           "classes": np.array(classes if classes is not None else [], np.int64),
           "counts": np.array([r.shape[0] for r in res], np.int64)}
    # the spy is only called for images that reach NMS; map calls back to images by order
    call = 0
    for b, r in enumerate(res):
        out[f"out_{b}"] = r.numpy().reshape(-1, 6)
        if r.shape[0]:
            boxes, scores, keep = rec[call]
            call += 1
            out[f"keep_{b}"] = keep[:300].numpy()
            out[f"ncand_{b}"] = np.array([boxes.shape[0]], np.int64)
        else:
            out[f"keep_{b}"] = np.zeros((0,), np.int64)
            # an image can reach NMS and still... no: NMS always keeps >= 1 box when n >= 1
            out[f"ncand_{b}"] = np.array([0], np.int64)
    assert call == len(rec)
    save(name, **out)


def synthetic_detections(B, N, C, seed, img=640.0, clusters=40, tie_every=0):
    """Decoded-detection-like matrix with heavy overlaps: boxes jitter around cluster centres."""
    g = torch.Generator()
    g.manual_seed(seed)
    cen = torch.rand(B, clusters, 2, generator=g) * img
    size = 20 + torch.rand(B, clusters, 2, generator=g) * 150
    which = torch.randint(0, clusters, (B, N), generator=g)
    c = torch.gather(cen, 1, which[..., None].expand(-1, -1, 2)) + torch.randn(B, N, 2, generator=g) * 6
    s = torch.gather(size, 1, which[..., None].expand(-1, -1, 2)) * (1 + 0.15 * torch.randn(B, N, 2, generator=g)).abs()
    box = torch.cat((c - s / 2, c + s / 2), -1)
    obj = torch.rand(B, N, 1, generator=g) ** 2
    cls = torch.rand(B, N, C, generator=g) ** 3
    if tie_every:
        obj[:, ::tie_every] = 0.5
        cls[:, ::tie_every] = 0.25
    return torch.cat((box, obj, cls), -1).float()


def nms_cases():
    det = run_decode("decode_s96_dense", 2, 96, 10, seed=43, obj_shift=4.0)
    run_nms("nms_decoded", det, 0.001, 0.6)
    run_nms("nms_clustered", synthetic_detections(3, 1200, 10, seed=51, tie_every=7), 0.001, 0.6)
    run_nms("nms_default_thr", synthetic_detections(2, 800, 10, seed=52), 0.25, 0.45)
    run_nms("nms_single_class", synthetic_detections(2, 900, 1, seed=53), 0.05, 0.5)
    run_nms("nms_class_filter", synthetic_detections(2, 500, 6, seed=54), 0.01, 0.6, classes=[1, 4])
    d = synthetic_detections(2, 400, 5, seed=55)
    d[1, :, 4] = 0.0  # image 1: nothing above the objectness threshold
    run_nms("nms_empty_image", d, 0.1, 0.6)
    # > 30000 candidates: exercises the top-max_nms cut (nms.py:61-63); scores are distinct w.h.p.
    big = synthetic_detections(1, 3600, 10, seed=56, clusters=400)
    big[..., 4] = 0.9 + 0.1 * big[..., 4]
    big[..., 5:] = 0.2 + 0.8 * big[..., 5:]
    run_nms("nms_over_max", big, 0.001, 0.6)


# ------------------------------------------------------------------------------------------------
# eval tail end to end: heads -> reference get_detections -> reference non_max_suppression
# ------------------------------------------------------------------------------------------------
def head_checksum(heads):
    """A few numbers that pin the regenerated logits to the ones the fixture was made from."""
    parts = [t.double().sum().item() for lv in heads for t in lv]
    parts += [float(lv[2].reshape(-1)[::9973][:8].double().sum()) for lv in heads]
    return np.array(parts, np.float64)


def run_evaltail(name, B, S, C, seed, obj_shift=0.0, inject=0.3, clutter=0, n_gt=None, conf=0.001, iou=0.6):
    """Heads are regenerated by the test from (B, S, C, seed, ...) with synthetic.make_head (same torch build in
    the build container and on the GPU box) and checked against `head_sums`."""
    heads = syn.make_head(B, S, C, seed=seed, obj_shift=obj_shift)
    targets = syn.make_targets(B, S, C, seed=seed, n_min=n_gt or 0, n_max=n_gt or 10)
    if clutter:
        syn.light_up_neighbourhoods(heads, targets, radius=clutter)
    elif inject > 0:
        syn.inject_ground_truth(heads, targets, S, fraction=inject, seed=seed)
    sums = head_checksum(heads)
    det = _Exp().get_detections(FeatureShape(width=S, height=S), net_result(heads))  # (clones: heads stay intact)
    rec = []
    real = torchvision.ops.nms

    def spy(boxes, scores, thr):
        keep = real(boxes, scores, thr)
        rec.append((boxes.clone(), scores.clone(), keep.clone()))
        return keep

    torchvision.ops.nms = spy
    try:
        res = ref_nms_mod.non_max_suppression(det.clone(), conf_thres=conf, nms_thres=iou)
    finally:
        torchvision.ops.nms = real
    out = {"meta": np.array([B, S, C, seed, clutter, n_gt or -1], np.int64),
           "fmeta": np.array([obj_shift, inject, conf, iou], np.float64), "head_sums": sums,
           "counts": np.array([r.shape[0] for r in res], np.int64)}
    keeps, keys, rows, ncand = [], [], [], []
    call = 0
    for b, r in enumerate(res):
        rows.append(r.numpy().reshape(-1, 6))
        if not r.shape[0]:
            ncand.append(0)
            continue
        boxes, scores, keep = rec[call]
        call += 1
        # identity (row, class) of every candidate handed to NMS: nms.py:33-63 re-run with torch ops
        x = det[b]
        xc = x[:, 4] > conf
        rid = xc.nonzero().squeeze(1)
        x = x[xc].clone()
        x[:, 5:] *= x[:, 4:5]
        if C > 1:
            i, j = (x[:, 5:] > conf).nonzero(as_tuple=False).T
            cs = x[i, j + 5]
        else:
            cs, j = x[:, 5:].max(1)
            m = cs > conf
            i, j, cs = m.nonzero().squeeze(1), j[m], cs[m]
        key = rid[i] * C + j
        n0 = int(key.shape[0])
        if n0 > 30000:
            order = cs.argsort(descending=True)[:30000]
            key, cs = key[order], cs[order]
        assert torch.equal(cs, scores), "candidate replay differs from what the reference handed to NMS"
        k = keep[:300]
        assert torch.equal(r[:, 4], cs[k])
        keeps.append(k.numpy())
        keys.append(key[k].numpy())
        ncand.append(n0)
    assert call == len(rec)
    out["keep_idx"] = np.concatenate(keeps) if keeps else np.zeros((0,), np.int64)
    out["keep_key"] = np.concatenate(keys) if keys else np.zeros((0,), np.int64)
    out["rows"] = np.concatenate(rows, 0).astype(np.float32)
    out["ncand"] = np.array(ncand, np.int64)
    save(name, **out)


def evaltail_cases():
    run_evaltail("evaltail_cfg1", 8, 640, 10, seed=2024)                 # BASELINE config[0]: B=8, 640, C=10
    run_evaltail("evaltail_cfg3", 24, 640, 10, seed=2026, inject=0.5)    # cfg 2/3 shape (per-image work is batch-independent)
    run_evaltail("evaltail_cfg5", 2, 1280, 80, seed=2028)                # cfg 5: > 30000 candidates, the top-k cut
    run_evaltail("evaltail_s416", 4, 416, 10, seed=2029, obj_shift=2.0)  # the reference's default input: 13x13 grid
    run_evaltail("evaltail_clutter", 2, 640, 10, seed=2030, clutter=2, n_gt=12)  # detector-like clutter


# ------------------------------------------------------------------------------------------------
# standalone IoUCalculator (iou.py:249-268), forward + autograd to both inputs
# ------------------------------------------------------------------------------------------------
def iou_cases():
    g = torch.Generator().manual_seed(61)
    n = 192
    c1 = torch.rand(n, 2, generator=g) * 100
    s1 = 4 + torch.rand(n, 2, generator=g) * 60
    # 2/3 of the pairs overlap (jittered copies), 1/3 are unrelated boxes (mostly disjoint)
    c2 = torch.where(torch.arange(n)[:, None] % 3 < 2, c1 + torch.randn(n, 2, generator=g) * 8,
                     torch.rand(n, 2, generator=g) * 100)
    s2 = s1 * (0.5 + torch.rand(n, 2, generator=g))
    b1 = torch.cat((c1 - s1 / 2, c1 + s1 / 2), 1).float()
    b2 = torch.cat((c2 - s2 / 2, c2 + s2 / 2), 1).float()
    up = torch.randn(n, generator=g).float()
    for it in ("iou", "giou", "diou", "ciou"):
        a = b1.clone().requires_grad_(True)
        b = b2.clone().requires_grad_(True)
        out = IoUCalculator(IoUType(it), eps=1e-7)(a, b)
        (out * up).sum().backward()
        save(f"iou_{it}", boxes1=b1.numpy(), boxes2=b2.numpy(), upstream=up.numpy(), out=out.detach().numpy(),
             grad1=a.grad.numpy(), grad2=b.grad.numpy(), iou_type=np.array([it]))


# ------------------------------------------------------------------------------------------------
# optimizer step + warm-up: SmartOptimizer(SGD nesterov) + OptimizerWarmupUpdater + sch_linear
# ------------------------------------------------------------------------------------------------
def _small_net(seed):
    import torch.nn as nn
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8), nn.SiLU(), nn.Conv2d(8, 5, 1, bias=True),
                         nn.GroupNorm(1, 5), nn.Flatten(), nn.LazyLinear(7))


def sgd_cases():
    from functools import partial
    from kod.nn.optim.smart import SmartOptimizer
    from kod.nn.optim.schedulers import sch_linear
    from kod.lightning.experiments.yv5_baseline.warmup import OptimizerWarmupUpdater
    net = _small_net(71)
    net(torch.zeros(1, 3, 6, 6))  # materialise the lazy layer
    names = [n for n, _ in net.named_parameters()]
    params = dict(net.named_parameters())
    opt = SmartOptimizer(partial(torch.optim.SGD, lr=0.01, momentum=0.937, nesterov=True), weight_decay=0.0005)(net)
    for pg in opt.param_groups:
        pg.setdefault("initial_lr", pg["lr"])  # what LambdaLR.__init__ records (schedulers.py:41-45)
    upd = OptimizerWarmupUpdater(warmup_epochs=3, warmup_bias_lr=0.1, warmup_momentum=0.8, momentum=0.937)
    sch = partial(sch_linear, max_epochs=30, lrf=0.01)
    out = {"names": np.array(names), "group_names": np.array([pg["name"] for pg in opt.param_groups])}
    group_of = {}
    for gi, pg in enumerate(opt.param_groups):
        for q in pg["params"]:
            group_of[id(q)] = gi
    out["group_index"] = np.array([group_of[id(params[n])] for n in names], np.int64)
    for n in names:
        out[f"init/{n}"] = params[n].detach().numpy().copy()
    g = torch.Generator().manual_seed(72)
    steps, nw = 7, 5  # warm-up for steps 0..5 (exp.py:166-185: global_step > nw returns early), plain SGD afterwards
    hyper = []
    for step in range(steps):
        for n in names:
            gr = torch.randn(params[n].shape, generator=g) * 0.1
            params[n].grad = gr.clone()
            out[f"grad{step}/{n}"] = gr.numpy()
        if step <= nw:
            upd(current_step=step, current_epoch=step // 3, max_warmup_steps=nw, sch_fn=sch, optimizer=opt)
        hyper.append([[float(pg["lr"]), float(pg["momentum"]), float(pg["weight_decay"])] for pg in opt.param_groups])
        opt.step()
    out["hyper"] = np.array(hyper, np.float64)
    for n in names:
        out[f"final/{n}"] = params[n].detach().numpy().copy()
        out[f"mom/{n}"] = opt.state[params[n]]["momentum_buffer"].numpy().copy()
    save("sgd_warmup", **out)


# ------------------------------------------------------------------------------------------------
# box transforms of the augmentation pipeline: mosaic, random perspective, flip, mixup
# ------------------------------------------------------------------------------------------------
def augment_cases():
    import random
    from kod.data.mosaic import MosaicAugmentor
    from kod.data.types import AugmentedSample
    from kod.data.augmentations import default as AD
    rng = np.random.default_rng(81)
    out = {}
    S = 320
    for case in range(4):
        tiles = []
        for i in range(4):
            h, w = int(rng.integers(120, 400)), int(rng.integers(120, 400))
            n = int(rng.integers(1, 9))
            c = rng.uniform(0, 1, (n, 2)) * [w, h]
            wh = rng.uniform(6, 160, (n, 2))
            b = np.concatenate((c - wh / 2, c + wh / 2), 1)
            b[:, [0, 2]] = b[:, [0, 2]].clip(0, w - 1)
            b[:, [1, 3]] = b[:, [1, 3]].clip(0, h - 1)
            tiles.append(AugmentedSample(image=np.zeros((h, w, 3), np.uint8), bboxes=b, labels=rng.integers(0, 10, n)))
        random.seed(900 + case)
        res, border = MosaicAugmentor(S)(tiles)
        random.seed(900 + case)  # replay the centre draw (mosaic.py:51-54)
        yc, xc = (int(random.uniform(-x, 2 * S + x)) for x in border)
        for i, t in enumerate(tiles):
            out[f"mosaic{case}/tile{i}_boxes"] = t.bboxes
            out[f"mosaic{case}/tile{i}_labels"] = t.labels
            out[f"mosaic{case}/tile{i}_wh"] = np.array([t.image.shape[1], t.image.shape[0]], np.int64)
        out[f"mosaic{case}/center"] = np.array([xc, yc], np.int64)
        out[f"mosaic{case}/out_boxes"] = res.bboxes
        out[f"mosaic{case}/out_labels"] = res.labels
        # the mosaic result goes through random_perspective with the mosaic border (default.py:279-351)
        vals = AD.AffineRandValues(perspective_x=0.0 if case % 2 == 0 else 2e-4, perspective_y=0.0 if case % 2 == 0 else -1e-4,
                                   degrees=float(rng.uniform(-10, 10)), scale=float(rng.uniform(0.6, 1.4)),
                                   shear_x=float(rng.uniform(-3, 3)), shear_y=float(rng.uniform(-3, 3)),
                                   translate_x=float(rng.uniform(0.4, 0.6)), translate_y=float(rng.uniform(0.4, 0.6)))
        persp = AD.random_perspective(res, vals, border=border)
        im = res.image
        feat = AD._get_feat_shape(width=im.shape[1], height=im.shape[0], border=border)
        M = AD._get_combined_matrix(affine_values=vals, feat_shape=FeatureShape(width=im.shape[1], height=im.shape[0]),
                                    feat_shape_with_border=feat, round_off=False)
        out[f"affine{case}/M"] = M
        out[f"affine{case}/out_wh"] = np.array([feat.width, feat.height], np.int64)
        out[f"affine{case}/perspective_scale"] = np.array([float(vals.perspective_x != 0 or vals.perspective_y != 0), vals.scale])
        out[f"affine{case}/out_boxes"] = persp.bboxes
        out[f"affine{case}/out_labels"] = persp.labels
        flipped = AD.horizontal_flip(AugmentedSample(np.zeros((feat.height, feat.width, 3), np.uint8), persp.bboxes, persp.labels))
        out[f"flip{case}/out_boxes"] = flipped.bboxes
        if case:
            prev = AugmentedSample(np.zeros((feat.height, feat.width, 3), np.float64), out[f"flip{case - 1}/out_boxes"],
                                   out[f"affine{case - 1}/out_labels"])
            cur = AugmentedSample(np.zeros((feat.height, feat.width, 3), np.float64), flipped.bboxes, flipped.labels)
            np.random.seed(3)
            mx = AD.mixup(prev, cur)
            out[f"mixup{case}/out_boxes"] = mx.bboxes
            out[f"mixup{case}/out_labels"] = mx.labels
    save("augment_boxes", **out)


# ------------------------------------------------------------------------------------------------
# Detect head: the three 1x1 convs + Rearrange of kod/nn/heads/yolov5.py, forward + backward
# ------------------------------------------------------------------------------------------------
def head_cases():
    from kod.nn.heads.yolov5 import Yolov5Head
    out = {}
    B, C, A = 2, 10, 3
    g = torch.Generator().manual_seed(91)
    # (in_channels, H, W, stride); 6x6 = 36 cells: ragged tile; 13x13 (the hl level of the reference's default 416 input)
    # and 48 channels: outside the tensor-core path (SIMT kernel)
    cfgs = [(64, 16, 16, 8), (128, 12, 12, 16), (256, 6, 6, 32), (128, 13, 13, 32), (48, 10, 7, 8)]
    out["meta"] = np.array([B, C, A], np.int64)
    out["cfgs"] = np.array(cfgs, np.int64)
    for l, (K, Hh, Ww, stride) in enumerate(cfgs):
        torch.manual_seed(92 + l)
        head = Yolov5Head(in_channels=K, num_anchors_per_cell=A, num_classes=C, stride=stride)
        x = torch.randn(B, K, Hh, Ww, generator=g).requires_grad_(True)
        r = head(x)
        ups = [torch.randn(t.shape, generator=g) for t in r]
        sum((t * u).sum() for t, u in zip(r, ups)).backward()
        out[f"l{l}/x"] = x.detach().numpy()
        out[f"l{l}/dx"] = x.grad.numpy()
        for name, p in head.state_dict().items():
            out[f"l{l}/param/{name}"] = p.numpy()
        for name, p in head.named_parameters():
            out[f"l{l}/grad/{name}"] = p.grad.numpy()
        for part, t, u in zip(("box", "obj", "cls"), r, ups):
            out[f"l{l}/{part}"] = t.detach().numpy()  # (B, A, H, W, P)
            out[f"l{l}/up_{part}"] = u.numpy()
    save("head_convs", **out)


def head_detect_cases():
    """Features -> reference Yolov5Head x3 -> reference get_detections (exp.py:70-102): the fused eval epilogue's target."""
    from kod.nn.heads.yolov5 import Yolov5Head
    B, C, A, S = 2, 10, 3, 128
    g = torch.Generator().manual_seed(191)
    out = {"meta": np.array([B, C, A, S], np.int64)}
    chans = (64, 128, 256)
    res = []
    for l, (K, stride) in enumerate(zip(chans, syn.STRIDES)):
        torch.manual_seed(192 + l)
        head = Yolov5Head(in_channels=K, num_anchors_per_cell=A, num_classes=C, stride=stride)
        x = torch.randn(B, K, S // stride, S // stride, generator=g)
        with torch.no_grad():
            r = head(x)
        res.append(r)
        out[f"l{l}/x"] = x.numpy()
        for name, p in head.state_dict().items():
            out[f"l{l}/param/{name}"] = p.numpy()
    with torch.no_grad():
        det = _Exp().get_detections(FeatureShape(width=S, height=S),
                                    Yolov5NetworkResult(*[DetectionHeadResult(*[t.clone() for t in r]) for r in res]))
    out["detections"] = det.numpy()
    save("head_detect", **out)


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "head":
        head_cases()
        head_detect_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "augment":
        augment_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sgd":
        sgd_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "iou":
        iou_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "evaltail":
        evaltail_cases()
        sys.exit(0)
    assign_cases()
    loss_cases()
    decode_cases()
    nms_cases()
    evaltail_cases()
    iou_cases()
    sgd_cases()
    augment_cases()
    head_cases()
    head_detect_cases()
