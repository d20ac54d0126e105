"""CPU: the oracle restatement against the golden vectors produced by the real reference."""
import numpy as np
import pytest
import torch

import _golden as G
from object_detection_cib_b200 import synthetic as syn
from oracle import assign as oa
from oracle import coco_map as om
from oracle import decode as od
from oracle import loss as ol
from oracle import nms as on

LEVELS = list(zip(syn.STRIDES, syn.ANCHORS_PX))


@pytest.mark.parametrize("name", G.names("assign_"))
def test_assign_matches_reference(name):
    g = G.load(name)
    res = oa.assign(G.targets_of(g), LEVELS, int(g["img_wh"][0]), int(g["img_wh"][1]), float(g["threshold"][0]))
    for ln, r in zip(G.LEVEL_NAMES, res):
        for key, val in (("samples", r.samples), ("anchors", r.anchors), ("grid_y", r.grid_y),
                         ("grid_x", r.grid_x), ("labels", r.labels)):
            assert np.array_equal(val, g[f"{ln}_{key}"]), (name, ln, key)
        assert np.array_equal(r.gt_boxes.view(np.uint32), g[f"{ln}_gt_boxes"].view(np.uint32)), (name, ln)
        assert np.array_equal(r.anchor_wh.view(np.uint32), g[f"{ln}_anchor_wh"].view(np.uint32)), (name, ln)


@pytest.mark.parametrize("name", G.names("loss_"))
def test_loss_matches_reference(name):
    g = G.load(name)
    B, S, C = (int(v) for v in g["meta"])
    heads = G.heads_of(g)
    targets = G.targets_of(g)
    asg = oa.assign(targets, LEVELS, S, S)
    hp = ol.LossHParams(*[float(v) for v in g["hparams"]])
    w = g["weights"].tolist() if g["weights"].size else None
    losses, grads = ol.loss_forward_backward(heads, asg, S, hp, str(g["iou_type"][0]), 1e-7, w,
                                             upstream=(B, B, B))
    got = np.array([float(v) for v in losses], np.float32)
    np.testing.assert_allclose(got, g["losses"], rtol=1e-6, atol=0, equal_nan=True)
    ref_grads = G.heads_of(g, grads=True)
    for lv, rlv in zip(grads, ref_grads):
        for a, b in zip(lv, rlv):
            np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", G.names("iou_"))
def test_iou_calculator_matches_reference(name):
    """oracle.loss.aligned_iou_xyxy against the real IoUCalculator (forward bit-exact, autograd to both inputs)."""
    g = G.load(name)
    a = torch.from_numpy(g["boxes1"]).requires_grad_(True)
    b = torch.from_numpy(g["boxes2"]).requires_grad_(True)
    out = ol.aligned_iou_xyxy(a, b, str(g["iou_type"][0]), 1e-7)
    assert np.array_equal(out.detach().numpy().view(np.uint32), g["out"].view(np.uint32))
    (out * torch.from_numpy(g["upstream"])).sum().backward()
    np.testing.assert_allclose(a.grad.numpy(), g["grad1"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(b.grad.numpy(), g["grad2"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", G.names("decode_"))
def test_decode_matches_reference(name):
    g = G.load(name)
    det = od.get_detections(G.heads_of(g), LEVELS)
    assert np.array_equal(det.numpy().view(np.uint32), g["detections"].view(np.uint32))


@pytest.mark.parametrize("name", G.names("nms_"))
def test_nms_matches_reference(name, oracle_lib):
    g = G.load(name)
    conf, iou = (float(v) for v in g["conf_iou"])
    classes = g["classes"].tolist() if g["classes"].size else None
    res = on.non_max_suppression(g["detections"], conf, iou, classes)
    assert [r.dets.shape[0] for r in res] == g["counts"].tolist()
    for b, r in enumerate(res):
        assert np.array_equal(r.keep, g[f"keep_{b}"]), (name, b)
        assert np.array_equal(r.dets.view(np.uint32), g[f"out_{b}"].view(np.uint32)), (name, b)
        if g[f"ncand_{b}"][0]:
            assert r.candidates.shape[0] == int(g[f"ncand_{b}"][0])


@pytest.mark.parametrize("name", G.names("evaltail_"))
def test_evaltail_matches_reference(name, oracle_lib):
    """heads -> oracle decode -> oracle NMS against the real reference run end to end (get_detections ->
    non_max_suppression on the CPU): counts, keep indices, candidate identities and rows, bit for bit.  A bounded
    number of images per fixture keeps the CPU suite short (the O(n^2) scan costs 1-2 s per dense image)."""
    g = G.load(name)
    heads, _, (B, S, C), (conf, iou) = G.evaltail_inputs(g)
    check = list(range(B))[:2] + ([B - 1] if B > 2 else [])
    # decode the WHOLE batch like the reference did: torch's CPU sigmoid is vectorised with a scalar tail, so the last
    # bit of a value depends on where the element sits in the tensor; a sliced batch would not reproduce it
    det = od.get_detections(heads, LEVELS).numpy()[check]
    res = on.non_max_suppression(det, conf, iou)
    ref = G.evaltail_split(g)
    for r, b in zip(res, check):
        kidx, kkey, rows = ref[b]
        assert r.dets.shape[0] == int(g["counts"][b]), (name, b)
        assert np.array_equal(r.keep, kidx), (name, b)
        got_rows, got_key = r.dets, r.cand_row[r.keep] * C + r.dets[:, 5].astype(np.int64)
        if r.n_before_cut > on.MAX_NMS:  # unstable argsort in the reference: exact ties in implementation-defined order
            got_rows, got_key = G.canon_ties(got_rows, got_key)
            rows, kkey = G.canon_ties(rows, kkey)
        assert np.array_equal(got_rows.view(np.uint32), rows.view(np.uint32)), (name, b)
        assert np.array_equal(got_key, kkey), (name, b)
        assert r.n_before_cut == int(g["ncand"][b])


def test_greedy_nms_against_torchvision(oracle_lib):
    """The C restatement against the third-party kernel it restates (torchvision.ops.nms, CPU)."""
    tv = pytest.importorskip("torchvision")
    gen = torch.Generator().manual_seed(5)
    for n, thr in ((1, 0.5), (57, 0.3), (700, 0.6), (1500, 0.45)):
        c = torch.rand(n, 2, generator=gen) * 100
        s = 5 + torch.rand(n, 2, generator=gen) * 40
        boxes = torch.cat((c - s / 2, c + s / 2), 1)
        scores = (torch.rand(n, generator=gen) * 16).round() / 16  # many exact ties
        want = tv.ops.nms(boxes, scores, thr).numpy()
        got = on.greedy_nms(boxes.numpy(), scores.numpy(), thr)
        assert np.array_equal(got, want)
    # fp32 ratio vs double threshold: 3/5 == (float)0.6 > 0.6 -> suppressed on the CPU
    b = np.array([[0, 0, 2, 2], [0.5, 0, 2.5, 2]], np.float32)
    assert on.greedy_nms(b, np.array([0.9, 0.8], np.float32), 0.6).tolist() == [0]


# ---- mAP restatement: parity unpinned; hand-computed known answers only -------------------------
def test_coco_map_known_answers(oracle_lib):
    gt_b = np.array([[0, 0, 10, 10], [20, 20, 40, 40]], np.float64)
    gt_l = np.array([0, 1])
    # perfect detections -> AP 1 at every threshold
    d = np.array([[0, 0, 10, 10, 0.9, 0], [20, 20, 40, 40, 0.8, 1]], np.float32)
    m = om.match_image(d, gt_b, gt_l, 3)
    assert m.tp.all() and m.gt_count.tolist() == [1, 1, 0]
    ap = om.average_precision(m.score, m.cls, m.tp, m.gt_count)
    assert np.allclose(ap[:, :2], 1.0) and np.isnan(ap[:, 2]).all()
    rep = om.report_from_ap(ap)
    assert abs(rep["avg_mAP"] - 1.0) < 1e-12 and sorted(rep["tag_wise_AP_50"]) == [0, 1]
    # one FP ranked above the TP: precision 1/2 at recall 1 -> AP = 0.5 (all 101 recall points)
    d = np.array([[50, 50, 60, 60, 0.95, 0], [0, 0, 10, 10, 0.9, 0]], np.float32)
    m = om.match_image(d, gt_b[:1], gt_l[:1], 1)
    assert m.tp[:, 1].tolist() == [False, True]
    ap = om.average_precision(m.score, m.cls, m.tp, m.gt_count)
    assert np.allclose(ap[:, 0], 0.5)
    # IoU exactly 0.5 counts at t=0.5 (>=), not at 0.75; box [0,0,10,5] vs GT [0,0,10,10]
    d = np.array([[0, 0, 10, 5, 0.9, 0]], np.float32)
    m = om.match_image(d, gt_b[:1], gt_l[:1], 1)
    assert m.tp[0].tolist() == [True, True, False, False]
    # equal IoU with two GTs: the later GT is taken, the earlier one stays free for the next detection
    g2 = np.array([[0, 0, 10, 10], [0, 0, 10, 10]], np.float64)
    d = np.array([[0, 0, 10, 10, 0.9, 0], [0, 0, 10, 10, 0.8, 0], [0, 0, 10, 10, 0.7, 0]], np.float32)
    m = om.match_image(d, g2, np.array([0, 0]), 1)
    assert m.tp[:, 1].tolist() == [True, True, False]
    # 2 GT, ranked TP FP TP: recall .5 -> p 1, recall 1 -> p 2/3: AP = (51*1 + 50*(2/3))/101
    gtb = np.array([[0, 0, 10, 10], [100, 100, 120, 120]], np.float64)
    d = np.array([[0, 0, 10, 10, 0.9, 0], [50, 50, 60, 60, 0.8, 0], [100, 100, 120, 120, 0.7, 0]], np.float32)
    m = om.match_image(d, gtb, np.array([0, 0]), 1)
    ap = om.average_precision(m.score, m.cls, m.tp, m.gt_count)
    assert abs(ap[1, 0] - (51 + 50 * (2 / 3)) / 101) < 1e-12


def test_coco_rule_kats():
    """Every pycocotools rule the restatement relies on, against answers derived by hand (tests/_coco_kats.py)."""
    import _coco_kats as K
    for case in K.CASES:
        C, ious = case["num_classes"], case["ious"]
        orc = om.CocoMapOracle(C, ious)
        ms = []
        for dets, gtb, gtl in case["images"]:
            m = om.match_image(dets, gtb, gtl, C, ious)
            ms.append(m)
            orc._matches.append(m)
        for (img, d), want in case["tp"].items():
            assert ms[img].tp[d].tolist() == want, (case["name"], img, d)
        s = np.concatenate([m.score for m in ms]); c = np.concatenate([m.cls for m in ms])
        tp = np.concatenate([m.tp for m in ms], 0); gcount = np.sum([m.gt_count for m in ms], 0)
        ap = om.average_precision(s, c, tp, gcount)
        for (t, k), want in case["ap"].items():
            if want is None:
                assert np.isnan(ap[t, k]), (case["name"], t, k)
            else:
                assert abs(ap[t, k] - want) < 1e-12, (case["name"], t, k, ap[t, k], want)
        rep = orc.get_report()
        if 0.5 in ious:
            assert sorted(rep["tag_wise_AP_50"]) == sorted(k for (t, k), w in case["ap"].items() if t == 1 and w is not None)


def test_sgd_warmup_matches_reference():
    """oracle.sgd against SmartOptimizer(SGD nesterov) + OptimizerWarmupUpdater run by the real reference classes:
    the interpolated lr / momentum per group and step, and parameters / momentum buffers after 7 steps, bit for bit."""
    from functools import partial
    from oracle import sgd as osgd
    g = G.load("sgd_warmup")
    names = [str(n) for n in g["names"]]
    gname = [str(n) for n in g["group_names"]]
    assert gname == ["bias_params", "decay_params", "norm_params"]  # smart.py:42-58
    gi = g["group_index"]
    assert [gname[i] for i in gi] == ["decay_params", "norm_params", "bias_params", "decay_params", "bias_params",
                                      "norm_params", "bias_params", "decay_params", "bias_params"]
    sch = partial(osgd.sch_linear, max_epochs=30, lrf=0.01)
    P = {n: g[f"init/{n}"] for n in names}
    M = {n: None for n in names}
    steps, nw = g["hyper"].shape[0], 5
    hyper = [[0.01, 0.937, wd] for wd in (0.0, 0.0005, 0.0)]
    for step in range(steps):
        if step <= nw:
            for k in range(3):
                hyper[k][0], hyper[k][1] = osgd.warmup_hyper(step, step // 3, nw, gname[k], 0.01, 0.1, 0.8, 0.937, sch)
        assert np.array_equal(np.array(hyper), g["hyper"][step]), step
        for n, k in zip(names, gi):
            P[n], M[n] = osgd.sgd_step(P[n], g[f"grad{step}/{n}"], M[n], *hyper[k])
    for n in names:
        assert np.array_equal(P[n].view(np.uint32), g[f"final/{n}"].view(np.uint32)), n
        assert np.array_equal(M[n].view(np.uint32), g[f"mom/{n}"].view(np.uint32)), n


def test_bbiou_c_matches_numpy(oracle_lib):
    rng = np.random.default_rng(3)
    d = rng.uniform(0, 50, (17, 4)); g = rng.uniform(0, 50, (9, 4))
    out = np.zeros((17, 9))
    on._lib().kod_oracle_bbiou(np.ascontiguousarray(d).ctypes.data, 17, np.ascontiguousarray(g).ctypes.data, 9,
                               out.ctypes.data)
    assert np.array_equal(out, om.bbiou(d, g))


def test_staged_reference_matches_the_port():
    """oracle/_ref (the reference's own files, staged by build()) and the oracle port give the same losses on the
    bench's CPU inputs: what `bench.py --impl reference` times is the path the parity tests pin."""
    import types

    import bench
    from oracle import ref_stage

    if not ref_stage.staged():
        pytest.skip("oracle/_ref not staged (no /root/reference on this box and no snapshot of it)")
    R = bench._ref_modules()
    assert R is not None and R["non_max_suppression"].__module__ == "kod.core.nms"
    a = types.SimpleNamespace(img=320, classes=10, batch=2)
    heads, targets = bench._cpu_inputs(a, 2)
    from object_detection_cib_b200 import synthetic as syn
    from oracle import assign as oa
    from oracle import loss as ol

    levels = list(zip(syn.STRIDES, syn.ANCHORS_PX))
    asg = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], levels, 320, 320)
    want, _ = ol.loss_forward_backward(heads, asg, 320, ol.LossHParams(), "ciou", 1e-7, syn.readme_class_weights(), upstream=(2, 2, 2))
    infos = [R["AnchorBoxInfo"](stride=s, boxes_wh=[R["FeatureShape"](width=w, height=h) for (w, h) in an]) for s, an in levels]
    mod = R["Yolov5Loss"](R["Yolov5LabelAssigner"](R["AssignmentAnchorInfo"](*infos)), R["Yolov5LossParams"].get_default(),
                          R["IoUCalculator"](R["IoUType"]("ciou"), eps=1e-7), weights=syn.readme_class_weights())
    mod.weights = mod.weights.cpu()  # (loss.py:58-61 moves them to cuda when one is visible)
    nr = R["Yolov5NetworkResult"](*[R["DetectionHeadResult"](*lv) for lv in heads])
    r = mod(R["FeatureShape"](width=320, height=320), nr, tuple(R["DetectionTarget"](boxes=b, labels=l) for b, l in targets))
    got = [float(r.localization), float(r.objectness), float(r.classification)]
    np.testing.assert_allclose(got, [float(v) for v in want], rtol=1e-6)
    bench._ref_step(a, R, heads, targets)  # the timed callable runs end to end
