"""GPU: the fused Detect head (kodb_head_forward: tcgen05 / TMEM / TMA tf32 GEMM with the planar epilogue)
against the real reference's Yolov5Head (three 1x1 convs + Rearrange, kod/nn/heads/yolov5.py:12-178).

Tolerance: the kernel multiplies in tf32 (10-bit mantissa, fp32 accumulate) -- the precision class of torch's default
convolution policy on this GPU (torch.backends.cudnn.allow_tf32 = True); the goldens are the reference's fp32 CPU
result.  |error| <= 2^-10 * sum_k |x_k w_k| in the worst case; with K <= 512 unit-variance inputs that is ~1e-2 absolute
at most and ~1e-3 typical: rtol 5e-3, atol 6e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import _golden as G
from object_detection_cib_b200.heads import Yolov5Head, fused_heads_detections, fused_heads_forward

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = dict(rtol=5e-3, atol=6e-3)


def _heads_from_golden(g):
    B, C, A = (int(v) for v in g["meta"])
    heads, xs = [], []
    for l, (K, Hh, Ww, stride) in enumerate(g["cfgs"].tolist()):
        h = Yolov5Head(K, A, C, stride).to(DEV)
        sd = {k[len(f"l{l}/param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f"l{l}/param/")}
        h.load_state_dict(sd)  # same parameter names as the reference module
        heads.append(h)
        xs.append(torch.from_numpy(g[f"l{l}/x"]).to(DEV).requires_grad_(True))
    return heads, xs, (B, C, A)


@pytest.mark.parametrize("one_launch", [True, False])
def test_head_golden_forward_backward(one_launch):
    g = G.load("head_convs")
    heads, xs, (B, C, A) = _heads_from_golden(g)
    # (one launch takes up to three levels: the five golden levels go in as 3 + 2)
    res = (fused_heads_forward(heads[:3], xs[:3]) + fused_heads_forward(heads[3:], xs[3:])) if one_launch else \
        [h(x) for h, x in zip(heads, xs)]
    loss = 0
    for l, r in enumerate(res):
        for part, t in zip(("box", "obj", "cls"), r):
            want = g[f"l{l}/{part}"]
            assert tuple(t.shape) == want.shape  # (B, A, H, W, P)
            # the physical layout is planar [B][A*P][H][W] (the view of Rearrange): cells contiguous
            assert t.stride(3) == 1 and t.stride(2) == t.shape[3] and t.stride(4) == t.shape[2] * t.shape[3]
            np.testing.assert_allclose(t.detach().cpu().numpy(), want, **TOL)
            loss = loss + (t * torch.from_numpy(g[f"l{l}/up_{part}"]).to(DEV)).sum()
    loss.backward()
    for l, (h, x) in enumerate(zip(heads, xs)):
        np.testing.assert_allclose(x.grad.cpu().numpy(), g[f"l{l}/dx"], rtol=5e-3, atol=2e-2)
        for name, p in h.named_parameters():
            want = g[f"l{l}/grad/{name}"]
            np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=5e-3, atol=5e-3 * max(1.0, float(np.abs(want).max())))


@pytest.mark.parametrize("B,S,C", [(4, 640, 10), (2, 320, 80), (3, 416, 10)])
def test_head_vs_torch_convs(B, S, C):
    """yolov5s channel widths (128 / 256 / 512) on full grids, against torch's own convolutions in strict fp32; a 416
    input has a 13x13 level (169 cells: not a multiple of 4) that the TMA path cannot take: that level runs on the SIMT
    kernel inside the same call."""
    A = 3
    torch.manual_seed(B + S + C)
    heads = [Yolov5Head(K, A, C, s).to(DEV) for K, s in ((128, 8), (256, 16), (512, 32))]
    xs = [torch.randn(B, K, S // s, S // s, device=DEV) for K, s in ((128, 8), (256, 16), (512, 32))]
    res = fused_heads_forward(heads, xs)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for h, x, r in zip(heads, xs, res):
            for m, t in zip((h.box_head, h.obj_head, h.cls_head), r):
                want = F.conv2d(x, m.conv.weight, m.conv.bias)  # (B, A*P, H, W)
                got = t.permute(0, 1, 4, 2, 3).reshape(want.shape)
                np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), **TOL)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("K,Hh,Ww,C", [(48, 10, 7, 10), (128, 13, 13, 3), (64, 8, 8, 91), (32, 1, 1, 1)])
def test_head_odd_shapes_simt_kernel(K, Hh, Ww, C):
    """Shapes outside the tensor-core path (in_channels % 32, H*W % 4, A*(5+C) > 256): same entry point, SIMT kernel, fp32
    arithmetic -- tight tolerance against torch's strict-fp32 convolutions; backward through the same autograd function."""
    A, B = 3, 3
    torch.manual_seed(K + C)
    h = Yolov5Head(K, A, C, 8).to(DEV)
    x = torch.randn(B, K, Hh, Ww, device=DEV, requires_grad=True)
    r = h(x)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for m, t in zip((h.box_head, h.obj_head, h.cls_head), r):
            want = F.conv2d(x, m.conv.weight, m.conv.bias)
            got = t.permute(0, 1, 4, 2, 3).reshape(want.shape)
            np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=1e-4, atol=1e-4)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    sum(t.sum() for t in r).backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()


def test_head_detections_golden():
    """Features -> detections in one launch (decode epilogue of the head GEMM) against the real reference's
    Yolov5Head x3 + get_detections (fp32 CPU).  tf32 logits (|err| ~ 2e-3) go through sigmoid / (2 sigma)^2 * anchor:
    probabilities to 1e-3, box corners to 5e-3 relative + 0.1 px."""
    from object_detection_cib_b200.types import FeatureShape, default_anchor_info
    g = G.load("head_detect")
    B, C, A, S = (int(v) for v in g["meta"])
    heads, xs = [], []
    for l, stride in enumerate((8, 16, 32)):
        x = torch.from_numpy(g[f"l{l}/x"]).to(DEV)
        h = Yolov5Head(x.shape[1], A, C, stride).to(DEV)
        h.load_state_dict({k[len(f"l{l}/param/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f"l{l}/param/")})
        heads.append(h)
        xs.append(x)
    det = fused_heads_detections(heads, xs, FeatureShape(S, S), default_anchor_info()).cpu().numpy()
    want = g["detections"]
    assert det.shape == want.shape
    np.testing.assert_allclose(det[..., :4], want[..., :4], rtol=5e-3, atol=0.1)
    np.testing.assert_allclose(det[..., 4:], want[..., 4:], rtol=0, atol=1e-3)


@pytest.mark.parametrize("B,S,C", [(4, 640, 10), (2, 320, 80), (3, 416, 10)])
def test_head_detections_vs_forward_then_decode(B, S, C):
    """The decode epilogue and the two-launch path (planar logits, then kodb_decode) read the same accumulators; the epilogue
    uses the approximate sigmoid (the logits are tf32 anyway), kodb_decode the exact one: probabilities agree to 1e-6
    relative, box corners to 1e-3 px (416: a 13x13 level -> the wrapper takes the two-launch path itself: identical)."""
    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200.types import FeatureShape, default_anchor_info
    torch.manual_seed(S + C)
    heads = [Yolov5Head(K, 3, C, s).to(DEV) for K, s in ((128, 8), (256, 16), (512, 32))]
    xs = [torch.randn(B, K, S // s, S // s, device=DEV) for K, s in ((128, 8), (256, 16), (512, 32))]
    info, shape = default_anchor_info(), FeatureShape(S, S)
    det = fused_heads_detections(heads, xs, shape, info)
    with torch.no_grad():
        want = N.get_detections(shape, fused_heads_forward(heads, xs), info)
    assert det.shape == want.shape == (B, sum(3 * (S // s) ** 2 for s in (8, 16, 32)), 5 + C)
    if S == 416:
        assert torch.equal(det, want)
    else:
        np.testing.assert_allclose(det[..., 4:].cpu().numpy(), want[..., 4:].cpu().numpy(), rtol=3e-6, atol=1e-9)
        np.testing.assert_allclose(det[..., :4].cpu().numpy(), want[..., :4].cpu().numpy(), rtol=1e-5, atol=1e-3)


def test_head_feeds_the_detection_tail():
    """The head's output IS the layout contract of the tail (a-0): loss and eval tail consume it without any copy."""
    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200 import synthetic as syn
    from object_detection_cib_b200.iou import IoUCalculator, IoUType
    from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
    from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams
    from object_detection_cib_b200.types import FeatureShape, default_anchor_info
    B, S, C = 2, 320, 10
    torch.manual_seed(5)
    heads = [Yolov5Head(K, 3, C, s).to(DEV) for K, s in ((128, 8), (256, 16), (512, 32))]
    xs = [torch.randn(B, K, S // s, S // s, device=DEV, requires_grad=True) for K, s in ((128, 8), (256, 16), (512, 32))]
    net_result = fused_heads_forward(heads, xs)
    info = default_anchor_info()
    loss = Yolov5Loss(Yolov5LabelAssigner(info), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou))
    targets = syn.make_targets(B, S, C, seed=3, n_min=1, n_max=6)
    r = loss(FeatureShape(S, S), net_result, targets)
    (B * (r.localization + r.objectness + r.classification)).backward()
    assert all(x.grad is not None and torch.isfinite(x.grad).all() for x in xs)
    assert all(p.grad is not None for h in heads for p in h.parameters())
    res = N.eval_tail(FeatureShape(S, S), [type(lv)(*[t.detach() for t in lv]) for lv in net_result], info, 0.001, 0.6)
    assert int(res.counts.min()) >= 0
