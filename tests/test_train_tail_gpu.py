"""GPU parity: CUDA target assignment + fused loss forward/backward, called through the C ABI
(ctypes -> libkodb.so), against the golden vectors of the real reference and against the oracle."""
import numpy as np
import pytest
import torch

import _golden as G
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.iou import IoUCalculator, IoUType
from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams
from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info
from oracle import assign as oa
from oracle import loss as ol

pytestmark = pytest.mark.gpu
LEVELS = list(zip(syn.STRIDES, syn.ANCHORS_PX))
DEV = "cuda:0"


def _tt(targets, device="cpu"):
    return [(torch.as_tensor(b).to(device), torch.as_tensor(l).to(device)) for b, l in targets]


def _check_assignment(res, want_levels):
    for r, w in zip(res, want_levels):
        got = dict(samples=r.indices.samples, anchors=r.indices.anchors, grid_y=r.indices.grid_y,
                   grid_x=r.indices.grid_x, labels=r.labels)
        for k, v in got.items():
            assert v.dtype == torch.long
            assert np.array_equal(v.cpu().numpy(), w[k]), k
        assert np.array_equal(r.gt_boxes.cpu().numpy().view(np.uint32), w["gt_boxes"].view(np.uint32))
        assert np.array_equal(r.anchors.cpu().numpy().view(np.uint32), w["anchor_wh"].view(np.uint32))


@pytest.mark.parametrize("name", G.names("assign_"))
@pytest.mark.parametrize("where", ["cpu", DEV])
def test_assign_golden(name, where):
    g = G.load(name)
    asg = Yolov5LabelAssigner(default_anchor_info(), threshold=float(g["threshold"][0]), device=DEV)
    res = asg(FeatureShape(int(g["img_wh"][0]), int(g["img_wh"][1])), _tt(G.targets_of(g), where))
    want = [{k: g[f"{ln}_{k}"] for k in ("samples", "anchors", "grid_y", "grid_x", "labels", "gt_boxes", "anchor_wh")}
            for ln in G.LEVEL_NAMES]
    _check_assignment(res, want)


@pytest.mark.parametrize("B,S,nmax,dtype", [(64, 640, 10, torch.float32), (16, 1280, 40, torch.float32),
                                            (7, 416, 72, torch.float64), (3, 96, 300, torch.float32)])
def test_assign_vs_oracle_large(B, S, nmax, dtype):
    targets = syn.make_targets(B, S, 10, seed=B * 7 + S, n_max=nmax, dtype=dtype, size_hi=S * 0.6)
    # centres snapped to cell boundaries -> duplicate rows
    b0, l0 = targets[0]
    if b0.shape[0]:
        c = ((b0[:, :2] + b0[:, 2:]) / 2 / 32).round() * 32
        targets[0] = (torch.cat((c - 20, c + 20), 1).clamp(0, S - 1).to(dtype), l0)
    res = Yolov5LabelAssigner(default_anchor_info(), device=DEV)(FeatureShape(S, S), targets)
    want = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], LEVELS, S, S)
    _check_assignment(res, [w._asdict() for w in want])


def _loss_module(hp, iou_type, weights):
    return Yolov5Loss(Yolov5LabelAssigner(default_anchor_info()), hp, IoUCalculator(IoUType(iou_type), 1e-7), weights)


def _run_cuda_loss(heads_cpu, targets, S, hp, iou_type, weights, upstream=None):
    leaves = []
    for lv in heads_cpu:
        leaves.append(DetectionHeadResult(*[t.to(DEV).detach().requires_grad_(True) for t in lv]))
    mod = _loss_module(hp, iou_type, weights)
    res = mod(FeatureShape(S, S), leaves, targets)
    B = heads_cpu[0][0].shape[0]
    u = upstream or (B, B, B)
    total = u[0] * res.localization + u[1] * res.objectness + u[2] * res.classification
    total.backward()
    losses = np.array([res.localization.item(), res.objectness.item(), res.classification.item()], np.float32)
    grads = [[t.grad.cpu() for t in lv] for lv in leaves]
    return losses, grads, mod


def _assert_grads_close(got, want, rtol=1e-5):
    for lv_g, lv_w in zip(got, want):
        for a, b in zip(lv_g, lv_w):
            a, b = a.numpy(), b.numpy()
            assert a.shape == b.shape
            scale = float(np.abs(b).max()) or 1.0
            # 1e-5 relative, with an absolute floor of 1e-5 x the tensor's largest gradient for
            # entries that are sums of cancelling terms
            np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("name", G.names("loss_"))
def test_loss_golden(name):
    g = G.load(name)
    B, S, C = (int(v) for v in g["meta"])
    hp = Yolov5LossParams(*[float(v) for v in g["hparams"]])
    w = g["weights"].tolist() if g["weights"].size else None
    losses, grads, _ = _run_cuda_loss(G.heads_of(g), _tt(G.targets_of(g)), S, hp, str(g["iou_type"][0]), w)
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5, atol=0, equal_nan=True)
    _assert_grads_close(grads, G.heads_of(g, grads=True))


@pytest.mark.parametrize("B,S,C,upstream", [(8, 640, 10, None), (4, 320, 80, (1.0, 1.0, 1.0)),
                                            (3, 416, 10, (2.0, 0.5, 3.0))])
def test_loss_vs_oracle(B, S, C, upstream):
    heads = syn.make_head(B, S, C, seed=S + C)
    targets = syn.make_targets(B, S, C, seed=S, n_max=10, empty_images=(1,))
    w = syn.readme_class_weights() if C == 10 else None
    hp = Yolov5LossParams.get_default()
    losses, grads, mod = _run_cuda_loss(heads, targets, S, hp, "ciou", w, upstream)
    asg = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], LEVELS, S, S)
    assert mod.last_num_assigned.tolist() == [a.samples.shape[0] for a in asg]
    want_l, want_g = ol.loss_forward_backward(heads, asg, S, ol.LossHParams(*hp), "ciou", 1e-7, w,
                                              upstream=upstream or (B, B, B))
    np.testing.assert_allclose(losses, np.array([float(v) for v in want_l], np.float32), rtol=1e-5, equal_nan=True)
    _assert_grads_close(grads, want_g)


@pytest.mark.parametrize("B,S,C,n_gt", [(2, 640, 10, 150), (3, 320, 80, 70)])
def test_loss_crowded_images(B, S, C, n_gt):
    """No limit on boxes per image (the reference has none): beyond ~58 boxes the per-(image, level) row records spill
    from shared memory to the workspace (and with 80 classes the class gradients are recomputed instead of stored).
    Same losses / gradients as the oracle; many boxes share cells, so the duplicate chains get long."""
    heads = syn.make_head(B, S, C, seed=S + n_gt)
    targets = syn.make_targets(B, S, C, seed=n_gt, n_min=n_gt, n_max=n_gt, size_lo=8, size_hi=120)
    hp = Yolov5LossParams.get_default()
    losses, grads, mod = _run_cuda_loss(heads, targets, S, hp, "ciou", None)
    asg = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], LEVELS, S, S)
    assert mod.last_num_assigned.tolist() == [a.samples.shape[0] for a in asg]
    want_l, want_g = ol.loss_forward_backward(heads, asg, S, ol.LossHParams(*hp), "ciou", 1e-7, None, upstream=(B, B, B))
    np.testing.assert_allclose(losses, np.array([float(v) for v in want_l], np.float32), rtol=1e-5)
    _assert_grads_close(grads, want_g)


def test_out_of_range_label_on_the_device_gives_nan_and_a_status_bit():
    """loss.py:145-152 raises IndexError for a class id outside [0, C).  Labels that are already on the device cannot be
    checked without a sync, so the kernel turns the losses of that launch into NaN and raises a status bit; the next
    launch with valid labels is clean again (the workspace re-arms itself)."""
    import ctypes
    from object_detection_cib_b200 import _lib as L
    B, S, C = 2, 128, 4
    heads = [DetectionHeadResult(*[t.to(DEV) for t in lv]) for lv in syn.make_head(B, S, C, seed=3)]
    good = _tt(syn.make_targets(B, S, C, seed=4, n_min=2, n_max=4, size_hi=100), DEV)
    bad = [(b.clone(), l.clone()) for b, l in good]
    bad[1][1][0] = C  # one label out of range
    mod = _loss_module(Yolov5LossParams.get_default(), "ciou", None)
    with torch.no_grad():
        r = mod(FeatureShape(S, S), heads, bad)
        plan = mod.plan(FeatureShape(S, S), heads, bad)
        assert all(bool(torch.isnan(v)) for v in r)
        st = ctypes.c_int32(-1)
        L.check(L.lib().kodb_train_status(plan.workspace.data_ptr(), ctypes.byref(st), None), "kodb_train_status")
        assert st.value & 1
        r2 = mod(FeatureShape(S, S), heads, good)
        assert not any(bool(torch.isnan(v)) for v in r2)
        L.check(L.lib().kodb_train_status(plan.workspace.data_ptr(), ctypes.byref(st), None), "kodb_train_status")
        assert st.value == 0
        # kodb_workspace_reset re-arms a workspace explicitly (recovery path after a failed launch)
        plan.workspace.fill_(0xAB)
        L.check(L.lib().kodb_workspace_reset(plan.workspace.data_ptr(), plan.workspace.numel(), None), "kodb_workspace_reset")
        r3 = mod(FeatureShape(S, S), heads, good)
        assert all(torch.equal(a, b) for a, b in zip(r2, r3))
    # host-side labels are rejected before the launch, like the reference
    with pytest.raises(IndexError):
        mod(FeatureShape(S, S), heads, [(b.cpu(), l.cpu()) for b, l in bad])


def test_loss_no_grad_and_determinism():
    B, S, C = 4, 256, 10
    heads = [DetectionHeadResult(*[t.to(DEV) for t in lv]) for lv in syn.make_head(B, S, C, seed=1)]
    targets = syn.make_targets(B, S, C, seed=2)
    mod = _loss_module(Yolov5LossParams.get_default(), "ciou", None)
    with torch.no_grad():
        a = mod(FeatureShape(S, S), heads, targets)
        b = mod(FeatureShape(S, S), heads, targets)
    assert all(torch.equal(x, y) for x, y in zip(a, b))  # fixed-order fp64 reduction
    assert not a.localization.requires_grad


def test_fails_loudly_on_cpu_tensors():
    heads = syn.make_head(1, 64, 3, seed=1)
    mod = _loss_module(Yolov5LossParams.get_default(), "ciou", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(FeatureShape(64, 64), heads, syn.make_targets(1, 64, 3))


def test_collated_targets_equal_the_tuple_form():
    """data.packed_collate_fn's HostPackedTargets (one pinned blob, one H2D) drives the loss and the eval tail to the
    same bits as the reference's tuple of DetectionTarget."""
    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200.data import HostPackedTargets
    from object_detection_cib_b200.map_eval import DeviceCocoMapEvaluator
    B, S, C = 5, 320, 10
    heads = syn.make_head(B, S, C, seed=21)
    targets = syn.make_targets(B, S, C, seed=21, n_min=0, n_max=9, empty_images=(3,))
    packed = HostPackedTargets.from_targets(targets).pin_memory()
    assert packed.blob.is_pinned()
    info, shape = default_anchor_info(), FeatureShape(S, S)
    loss = Yolov5Loss(Yolov5LabelAssigner(info), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou), syn.readme_class_weights())
    out = []
    for tg in (targets, packed):
        leaves = [DetectionHeadResult(*[t.to(DEV).requires_grad_(True) for t in lv]) for lv in heads]
        r = loss(shape, leaves, tg)
        (r.localization + r.classification + r.objectness).backward()
        ev = DeviceCocoMapEvaluator(C, device=DEV)
        res, rec = N.eval_tail(shape, [DetectionHeadResult(*[t.detach() for t in lv]) for lv in leaves], info, 0.001, 0.6,
                               targets=tg, gt_count=ev.gt_count)
        out.append((torch.stack(list(r)).detach(), [t.grad for lv in leaves for t in lv], rec.tp.clone(), ev.gt_count.clone()))
    assert torch.equal(out[0][0], out[1][0])
    for ga, gb in zip(out[0][1], out[1][1]):
        assert torch.equal(ga, gb)
    assert torch.equal(out[0][2], out[1][2]) and torch.equal(out[0][3], out[1][3])


# ---- standalone IoUCalculator (iou.py:249-268) ------------------------------------------------------------------------
@pytest.mark.parametrize("name", G.names("iou_"))
def test_iou_calculator_golden(name):
    """IoUCalculator.__call__ on the CUDA kernel against the real reference: values to fp32 rounding (atan, division),
    gradients to both inputs at rtol 1e-5."""
    from object_detection_cib_b200.iou import IoUCalculator, IoUType
    g = G.load(name)
    a = torch.from_numpy(g["boxes1"]).to(DEV).requires_grad_(True)
    b = torch.from_numpy(g["boxes2"]).to(DEV).requires_grad_(True)
    out = IoUCalculator(IoUType(str(g["iou_type"][0])), 1e-7)(a, b)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], rtol=2e-6, atol=2e-7)
    (out * torch.from_numpy(g["upstream"]).to(DEV)).sum().backward()
    for got, want in ((a.grad, g["grad1"]), (b.grad, g["grad2"])):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("it", ["iou", "giou", "diou", "ciou"])
def test_iou_calculator_vs_oracle_batched(it):
    """(B, N, 4) inputs, one input without grad, forward-only call."""
    from object_detection_cib_b200.iou import IoUCalculator, IoUType
    gen = torch.Generator().manual_seed(17)
    c = torch.rand(3, 4097, 2, generator=gen) * 300
    s = 2 + torch.rand(3, 4097, 2, generator=gen) * 90
    b1 = torch.cat((c - s / 2, c + s / 2), -1)
    c2 = c + torch.randn(3, 4097, 2, generator=gen) * 15
    s2 = s * (0.6 + 0.8 * torch.rand(3, 4097, 2, generator=gen))
    b2 = torch.cat((c2 - s2 / 2, c2 + s2 / 2), -1)
    want_in = b1.clone().requires_grad_(True)
    want = ol.aligned_iou_xyxy(want_in, b2, it, 1e-7)
    want.sum().backward()
    calc = IoUCalculator(IoUType(it))
    a = b1.to(DEV).requires_grad_(True)
    got = calc(a, b2.to(DEV))
    assert got.shape == (3, 4097)
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().numpy(), rtol=2e-6, atol=2e-7)
    got.sum().backward()
    wg = want_in.grad.numpy()
    np.testing.assert_allclose(a.grad.cpu().numpy(), wg, rtol=1e-5, atol=1e-5 * np.abs(wg).max())
    with torch.no_grad():
        assert torch.equal(calc(b1.to(DEV), b2.to(DEV)), got.detach())
    with pytest.raises(RuntimeError):
        calc(b1, b2)  # CPU tensors: no fallback
