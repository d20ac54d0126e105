"""GPU: the Lightning-facing glue -- `Yolov5TailMixin` (exp.py:104-154) and `PyCOCOMAPEvalCallback`
(pycoco_map_eval.py:41-144) -- driven the way the Trainer drives the reference classes, with the attribute names the
reference experiment has, and checked against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from object_detection_cib_b200 import nms as N
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.data import HostPackedTargets
from object_detection_cib_b200.experiment import Yolov5TailMixin
from object_detection_cib_b200.iou import IoUCalculator, IoUType
from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams
from object_detection_cib_b200.map_eval import DeviceCocoMapEvaluator, PyCOCOMAPEvalCallback
from object_detection_cib_b200.types import DetectionHeadResult, Yolov5NetworkResult, default_anchor_info
from oracle import assign as oa
from oracle import coco_map as om
from oracle import loss as ol
from oracle import nms as on

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LEVELS = list(zip(syn.STRIDES, syn.ANCHORS_PX))
LABELS = {i: n for i, n in enumerate(("person", "car", "dining table", "chair", "cup", "bottle", "cat", "dog", "truck", "sink"))}


class _LightningStub(nn.Module):
    """What the mixin and the callback touch of LightningModule: log / log_dict (recorded)."""

    def __init__(self):
        super().__init__()
        self.logged, self.logged_dicts = {}, []

    def log(self, name, value, **kw):
        self.logged[name] = (value, kw)

    def log_dict(self, d, **kw):
        self.logged_dicts.append((dict(d), kw))


class _FixedHeadNet(nn.Module):
    """Stands in for Yolov5Network: returns head maps that are leaves of the graph (like conv outputs would be)."""

    def __init__(self, heads):
        super().__init__()
        self.heads = heads

    def forward(self, images):
        return Yolov5NetworkResult(*[DetectionHeadResult(*lv) for lv in self.heads])


class _Trainer:
    current_epoch = 3


def _experiment(B, S, C, seed, requires_grad):
    heads = syn.make_head(B, S, C, seed=seed, obj_shift=1.0)
    targets = syn.make_targets(B, S, C, seed=seed, n_min=0, n_max=9, empty_images=(1,))
    if not requires_grad:
        # eval side only: a prediction that coincides with its target sits on the kink of IoU's max/min, where a
        # one-ulp difference flips a gradient's sign (measure-zero set)
        syn.inject_ground_truth(heads, targets, S, fraction=0.6, seed=seed)

    class Experiment(Yolov5TailMixin, _LightningStub):  # the MRO INTEGRATION.md prescribes
        pass

    e = Experiment()
    dev_heads = [[t.to(DEV).requires_grad_(requires_grad) for t in lv] for lv in heads]
    e.net = _FixedHeadNet(dev_heads)
    e.anchor_info = default_anchor_info()
    e.loss = Yolov5Loss(Yolov5LabelAssigner(e.anchor_info), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou),
                        syn.readme_class_weights())
    e.val_nms_conf_threshold, e.val_nms_iou_threshold = 0.001, 0.6  # exp.py:45-46
    images = torch.zeros(B, 3, S, S, device=DEV)
    return e, heads, dev_heads, targets, images


def test_training_step_matches_reference_semantics():
    B, S, C = 4, 320, 10
    e, heads, dev_heads, targets, images = _experiment(B, S, C, 5, True)
    batch = (images, tuple(targets), None)
    total = e.training_step(batch, 0)
    total.backward()
    asg = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], LEVELS, S, S)
    (loc, obj, cls), grads = ol.loss_forward_backward(heads, asg, S, ol.LossHParams(), "ciou", 1e-7,
                                                      syn.readme_class_weights(), upstream=(B, B, B))
    # exp.py:124-130: batch_size * (loc + cls + obj); the three parts are logged as box / cls / obj with the batch size
    np.testing.assert_allclose(float(total.detach()), B * (float(loc) + float(cls) + float(obj)), rtol=1e-5)
    assert set(e.logged) == {"obj", "cls", "box"}
    for name, want in (("box", loc), ("obj", obj), ("cls", cls)):
        val, kw = e.logged[name]
        np.testing.assert_allclose(float(val), float(want), rtol=1e-5)
        assert kw == {"prog_bar": True, "batch_size": B}
    for lv, wl in zip(dev_heads, grads):
        for t, g in zip(lv, wl):
            scale = float(g.abs().max()) or 1.0
            np.testing.assert_allclose(t.grad.cpu().numpy(), g.numpy(), rtol=1e-5, atol=1e-5 * scale)


def _oracle_report(e, dev_heads, targets, S, C):
    from object_detection_cib_b200.types import FeatureShape
    det = N.get_detections(FeatureShape(S, S), e.net(None), e.anchor_info).cpu().numpy()
    orc = om.CocoMapOracle(C)
    ref = on.non_max_suppression(det, 0.001, 0.6)
    for r, (gb, gl) in zip(ref, targets):
        orc.add_image(r.dets, gb.numpy(), gl.numpy())
    return ref, orc.get_report()


@pytest.mark.parametrize("mode", ["reference_list", "padded", "fused_records", "collated_targets"])
def test_validation_step_and_map_callback(mode):
    """validation_step -> on_validation_batch_end -> on_validation_epoch_end over two batches: the logged keys are the
    reference's (map, map30, map50, map75, map90, map50_<class>) and the values equal the oracle's report, for the
    reference's output format (list of (n,6) tensors), the padded no-sync format and the fused-matching path."""
    B, S, C = 3, 320, 10
    cb = PyCOCOMAPEvalCallback(LABELS, device=DEV)
    trainer = _Trainer()
    orc = om.CocoMapOracle(C)
    e = None
    for step, seed in enumerate((11, 12)):
        e, heads, dev_heads, targets, images = _experiment(B, S, C, seed, False)
        if mode == "padded":
            e.tail_detections_format = "padded"
        if mode == "fused_records":
            e.tail_evaluator = cb.evaluator
        tg = HostPackedTargets.from_targets(targets).to(DEV) if mode == "collated_targets" else tuple(targets)
        out = e.validation_step((images, tg, None), step)
        assert set(out) >= {"targets", "detections"}  # MAPEvalCallbackInput (callbacks/map_eval.py:13-15)
        ref, _ = _oracle_report(e, dev_heads, targets, S, C)
        if mode in ("reference_list", "fused_records", "collated_targets"):
            assert isinstance(out["detections"], list) and len(out["detections"]) == B
            for d, r in zip(out["detections"], ref):
                assert np.array_equal(d.cpu().numpy().view(np.uint32), r.dets.view(np.uint32))
        for r, (gb, gl) in zip(ref, targets):
            orc.add_image(r.dets, gb.numpy(), gl.numpy())
        cb.on_validation_batch_end(trainer, e, out, (images, tg, None), step)
    want = orc.get_report()
    cb.on_validation_epoch_end(trainer, e)
    assert len(e.logged_dicts) == 1
    logged, kw = e.logged_dicts[0]
    expect_keys = {"map", "map30", "map50", "map75", "map90"} | {f"map50_{LABELS[k]}" for k in want["tag_wise_AP_50"]}
    assert set(logged) == expect_keys
    for k_log, k_rep in (("map", "avg_mAP"), ("map30", "mAP_30"), ("map50", "mAP_50"), ("map75", "mAP_75"), ("map90", "mAP_90")):
        assert abs(logged[k_log] - want[k_rep]) < 1e-4, (k_log, logged[k_log], want[k_rep])
    for k, v in want["tag_wise_AP_50"].items():
        assert abs(logged[f"map50_{LABELS[k]}"] - v) < 1e-4
    assert want["mAP_50"] > 0.05, "the synthetic set must contain true positives"
    # reset() at epoch end (pycoco_map_eval.py:144): a new epoch starts from nothing
    assert cb.evaluator._records == [] and int(cb.evaluator.gt_count.sum()) == 0


def test_get_detections_override_matches_oracle():
    from object_detection_cib_b200.types import FeatureShape
    from oracle import decode as od
    B, S, C = 2, 416, 10
    e, heads, dev_heads, targets, images = _experiment(B, S, C, 9, False)
    det = e.get_detections(FeatureShape(S, S), e.net(images))
    want = od.get_detections([tuple(lv) for lv in dev_heads], LEVELS)  # the reference's ops on the GPU
    assert torch.equal(det.view(torch.int32), want.view(torch.int32))
