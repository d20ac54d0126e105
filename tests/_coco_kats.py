"""Hand-derived known answers for every pycocotools rule the COCO restatement (oracle/coco_map.py) and the CUDA
matching / AP kernels rely on.  pycocotools / vision_evaluation are absent from the image, so this stage has no
reference-generated golden ("parity unpinned"); each case below states the rule, the inputs and the answer worked
out by hand from the published algorithm (COCOeval.evaluateImg / accumulate, maskApi.c::bbIou).

A case = dict(name, num_classes, ious, images=[(dets (n,6) xyxy+score+cls, gt_boxes (g,4), gt_labels (g,))],
              tp={(image, det): [bool per threshold]}, ap={(threshold index, class): value or None for "no GT"}).
"""
import numpy as np

T = (0.3, 0.5, 0.75, 0.9)


def _img(dets, gtb, gtl):
    return (np.array(dets, np.float32).reshape(-1, 6), np.array(gtb, np.float64).reshape(-1, 4),
            np.array(gtl, np.int64).reshape(-1))


CASES = [
    # (1) IoU exactly equal to the threshold matches (`iou < best: continue` with best = t).  Box [0,0,10,5] in GT
    #     [0,0,10,10]: i = 50, u = 100 -> IoU 0.5 exactly.  TP at 0.3 and 0.5, FP at 0.75 / 0.9.
    dict(name="iou_equals_threshold", num_classes=1, ious=T,
         images=[_img([[0, 0, 10, 5, 0.9, 0]], [[0, 0, 10, 10]], [0])],
         tp={(0, 0): [True, True, False, False]},
         ap={(0, 0): 1.0, (1, 0): 1.0, (2, 0): 0.0, (3, 0): 0.0}),
    # (2) equal IoU with two free ground truths: the LATER one is taken, the earlier one stays free.  Three identical
    #     detections on two identical GT boxes: det 0 takes GT 1, det 1 takes GT 0, det 2 finds nothing.
    #     AP: tp = [1,1,0], recall (.5, 1, 1), precision (1, 1, 2/3) -> every recall point reads 1.0.
    dict(name="equal_iou_later_gt_wins", num_classes=1, ious=T,
         images=[_img([[0, 0, 10, 10, 0.9, 0], [0, 0, 10, 10, 0.8, 0], [0, 0, 10, 10, 0.7, 0]],
                      [[0, 0, 10, 10], [0, 0, 10, 10]], [0, 0])],
         tp={(0, 0): [True] * 4, (0, 1): [True] * 4, (0, 2): [False] * 4},
         ap={(t, 0): 1.0 for t in range(4)}),
    # (3) the larger IoU wins over the later ground truth: GT 0 overlaps the detection fully, GT 1 only by half.
    dict(name="larger_iou_wins", num_classes=1, ious=T,
         images=[_img([[0, 0, 10, 10, 0.9, 0]], [[0, 0, 10, 10], [0, 0, 10, 20]], [0, 0])],
         tp={(0, 0): [True] * 4},
         # one of two GT found by the only detection: recall .5, precision 1 -> 51 of 101 recall points read 1
         ap={(t, 0): 51 / 101 for t in range(4)}),
    # (4) threshold 1.0 is capped at 1 - 1e-10: a perfect detection still matches at t = 1.0.
    dict(name="threshold_cap", num_classes=1, ious=(1.0, 0.5, 0.75, 0.9),
         images=[_img([[3, 4, 13, 24, 0.9, 0]], [[3, 4, 13, 24]], [0])],
         tp={(0, 0): [True] * 4}, ap={(t, 0): 1.0 for t in range(4)}),
    # (5) score ties across images keep image order (mergesort is stable).  Image 0: FP with score 0.5; image 1: TP with
    #     score 0.5; two GT in total.  Order [FP, TP]: recall (0, .5), precision (0, .5) -> monotone (.5, .5); recall
    #     points 0..0.5 (51 of them) read 0.5, the rest 0: AP = 51 * 0.5 / 101.  (The other order would give 51/101.)
    dict(name="score_ties_across_images", num_classes=1, ious=T,
         images=[_img([[50, 50, 60, 60, 0.5, 0]], [[0, 0, 10, 10]], [0]),
                 _img([[0, 0, 10, 10, 0.5, 0]], [[0, 0, 10, 10]], [0])],
         tp={(0, 0): [False] * 4, (1, 0): [True] * 4},
         ap={(t, 0): 51 * 0.5 / 101 for t in range(4)}),
    # (6) a class without any ground truth is excluded from every mean (precision = -1 in pycocotools) and from the
    #     tag-wise table; its detections do not hurt the other classes.
    dict(name="class_without_gt", num_classes=3, ious=T,
         images=[_img([[0, 0, 10, 10, 0.9, 0], [0, 0, 10, 10, 0.95, 2]], [[0, 0, 10, 10]], [0])],
         tp={(0, 0): [True] * 4, (0, 1): [False] * 4},
         ap={(1, 0): 1.0, (1, 1): None, (1, 2): None}),
    # (7) a detection of a class that has ground truth only in ANOTHER image is a false positive of that class.
    #     Class 1: image 0 has the GT and no detection, image 1 has a detection (score .9) and no GT of class 1 -> 1 FP,
    #     0 TP, recall never leaves 0 -> AP 0.  Class 0 is perfect.
    dict(name="fp_from_image_without_that_class", num_classes=2, ious=T,
         images=[_img([[0, 0, 10, 10, 0.8, 0]], [[0, 0, 10, 10], [20, 20, 30, 30]], [0, 1]),
                 _img([[20, 20, 30, 30, 0.9, 1]], [[5, 5, 9, 9]], [0])],
         tp={(0, 0): [True] * 4, (1, 0): [False] * 4},
         # class 0: two GT (one per image), one found: AP 51/101; class 1: AP 0
         ap={(1, 0): 51 / 101, (1, 1): 0.0}),
    # (8) TP, FP, TP over two ground truths: recall (.5, .5, 1), precision (1, .5, 2/3) -> monotone (1, 2/3, 2/3);
    #     recall points 0..0.5 read 1 (51 points), 0.51..1 read 2/3 (50 points).
    dict(name="tp_fp_tp", num_classes=1, ious=T,
         images=[_img([[0, 0, 10, 10, 0.9, 0], [50, 50, 60, 60, 0.8, 0], [100, 100, 120, 120, 0.7, 0]],
                      [[0, 0, 10, 10], [100, 100, 120, 120]], [0, 0])],
         tp={(0, 0): [True] * 4, (0, 1): [False] * 4, (0, 2): [True] * 4},
         ap={(t, 0): (51 + 50 * (2 / 3)) / 101 for t in range(4)}),
    # (9) a matched ground truth is not available to a later, better-overlapping detection (greedy in score order).
    #     det 0 (score .9) overlaps the GT with IoU 0.5 and takes it at t <= 0.5; det 1 (score .8) is the perfect box:
    #     FP at 0.3 / 0.5 (GT taken), TP at 0.75 / 0.9 (det 0 failed there, GT still free).
    dict(name="greedy_in_score_order", num_classes=1, ious=T,
         images=[_img([[0, 0, 10, 5, 0.9, 0], [0, 0, 10, 10, 0.8, 0]], [[0, 0, 10, 10]], [0])],
         tp={(0, 0): [True, True, False, False], (0, 1): [False, False, True, True]},
         # t<=0.5: [TP, FP] -> AP 1;  t>=0.75: [FP, TP] -> precision (0, .5) -> monotone .5 at every recall point
         ap={(0, 0): 1.0, (1, 0): 1.0, (2, 0): 0.5, (3, 0): 0.5}),
    # (10) touching boxes (w <= 0 or h <= 0 in bbIou) have IoU 0 and never match, even at a tiny threshold.
    dict(name="touching_boxes", num_classes=1, ious=(1e-9, 0.5, 0.75, 0.9),
         images=[_img([[10, 0, 20, 10, 0.9, 0]], [[0, 0, 10, 10]], [0])],
         tp={(0, 0): [False] * 4}, ap={(t, 0): 0.0 for t in range(4)}),
]
