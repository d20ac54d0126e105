"""CPU oracle for the YOLOv5 detection-tail hot path of craston/object_detection_cib (`kod`).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, on the CPU, the arithmetic of the reference functions listed in SURVEY.md §8(a):

* `oracle.assign`   – `Yolov5LabelAssigner` (kod/core/label_assignment/yv5.py:45-319)
* `oracle.loss`     – `Yolov5Loss` + `compute_{iou,giou,diou,ciou}`
                      (kod/lightning/experiments/yv5_baseline/loss.py:46-248, kod/core/bbox/iou.py:77-246)
* `oracle.decode`   – `get_detections` / `Yolov5Prediction` (exp.py:70-102, layers.py:15-155)
* `oracle.nms`      – `non_max_suppression` (kod/core/nms.py:9-75) + torchvision.ops.nms (third party)
* `oracle.coco_map` – pycocotools `COCOeval` matching/accumulate as driven by
                      `PyCOCOMAPEvalCallback` (kod/lightning/callbacks/pycoco_map_eval.py:41-144)

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it, and only as the checker (or the timed CPU baseline) – the product
package `object_detection_cib_b200` never imports `oracle` and has no CPU fallback.

Pinning status (see DESIGN.md §"Oracle"):
* assign / loss / decode / nms: pinned against outputs of the *real* reference executed in the
  build container (`tests/golden/make_golden.py` → `tests/golden/*.npz`).
* coco_map: **parity unpinned** – `vision_evaluation` / `pycocotools` are neither under
  `/root/reference` nor installable here; the restatement follows pycocotools 2.0.x semantics
   and is checked only against hand-computed known answers.
"""
