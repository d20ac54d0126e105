"""mAP evaluation drop-ins: the COCO matching + accumulation the reference obtains from
`vision_evaluation.evaluators.CocoMeanAveragePrecisionEvaluator` (-> pycocotools) inside
`PyCOCOMAPEvalCallback` (kod/lightning/callbacks/pycoco_map_eval.py:41-144), on the CUDA kernels
`kodb_map_match` / the fused eval tail (csrc/eval.cu) and `kodb_ap_accumulate` (csrc/ap.cu).

`DeviceCocoMapEvaluator` keeps the evaluator protocol the callback uses – `add_predictions(preds, gts)`,
`get_report()`, `reset()` – and adds device-side entry points that avoid the reference's per-row
`.cpu()` / `.tolist()` loops (pycoco_map_eval.py:66-104).  Records stay on the GPU for the whole
validation epoch; the report needs one D2H copy of the (4, C) AP table.
"""
from __future__ import annotations

from typing import Any, Optional, Sequence

import torch

from . import _host as H
from . import _lib as L
from . import distributed as D
from .nms import IOU_THRESHOLDS, MAX_DET, MatchRecords, NmsResult, _alloc_match

try:  # lightning is optional (absent in the build image); the callback only needs the hook names
    from lightning.pytorch.callbacks import Callback as _CallbackBase
except Exception:  # pragma: no cover
    _CallbackBase = object


class DeviceCocoMapEvaluator(object):
    def __init__(self, num_classes: int, ious: Sequence[float] = IOU_THRESHOLDS,
                 report_tag_wise: Sequence[bool] = (False, True, False, False), max_det: int = MAX_DET,
                 device="cuda"):
        self.num_classes = int(num_classes)
        self.ious = tuple(float(t) for t in ious)
        self.report_tag_wise = tuple(bool(v) for v in report_tag_wise)
        self.max_det = int(max_det)
        self.device = torch.device(device)
        self.reset()

    # ---- state ------------------------------------------------------------------------------------
    def reset(self) -> None:
        self._records: list[MatchRecords] = []
        self.gt_count = torch.zeros(self.num_classes, dtype=torch.long, device=self.device)

    # ---- feeding ----------------------------------------------------------------------------------
    def add_records(self, records: MatchRecords) -> None:
        """Records produced by the fused eval tail (`nms.eval_tail(..., targets=, gt_count=self.gt_count)`)."""
        self._records.append(records)

    def add_detections(self, dets: torch.Tensor, counts: torch.Tensor, targets) -> None:
        """Padded NMS output (B, max_det, 6) + counts + the batch's targets -> matching on the GPU."""
        H.require_cuda(dets, "detections")
        B = dets.shape[0]
        dev = dets.device
        packed = targets if isinstance(targets, H.PackedTargets) else H.pack_targets(targets, dev)
        counts = counts.to(device=dev, dtype=torch.int32)
        m = _alloc_match(B, self.max_det, self.gt_count, counts, dev, self.ious)
        L.check(L.lib().kodb_map_match(dets.contiguous().data_ptr(), counts.data_ptr(), B, self.max_det,
                                       self.num_classes, packed.struct, m.struct, H.stream_ptr(dev)), "kodb_map_match")
        self._records.append(m.records)

    def add_detection_list(self, detections: Sequence[torch.Tensor], targets) -> None:
        """The reference's validation_step output: a list of (n_i, 6) tensors (exp.py:154)."""
        B = len(detections)
        dets = torch.zeros((B, self.max_det, 6), dtype=torch.float32, device=self.device)
        counts = []
        for b, d in enumerate(detections):
            n = min(int(d.shape[0]), self.max_det)
            counts.append(n)
            if n:
                dets[b, :n] = d[:n].to(self.device, torch.float32)
        self.add_detections(dets, torch.tensor(counts, dtype=torch.int32), targets)

    def add_predictions(self, predictions, targets) -> None:
        """vision_evaluation protocol (pycoco_map_eval.py:106): python lists of
        (label, score, l, t, r, b) per image and (label, l, t, r, b) per image."""
        dets = []
        for img in predictions:
            rows = [[p[2], p[3], p[4], p[5], p[1], p[0]] for p in img]
            dets.append(torch.tensor(rows, dtype=torch.float32).reshape(-1, 6))
        tg = []
        for img in targets:
            boxes = torch.tensor([[g[1], g[2], g[3], g[4]] for g in img], dtype=torch.float64).reshape(-1, 4)
            labels = torch.tensor([int(g[0]) for g in img], dtype=torch.long)
            tg.append((boxes, labels))
        self.add_detection_list(dets, tg)

    # ---- report -----------------------------------------------------------------------------------
    def average_precision(self, gather: bool = True) -> torch.Tensor:
        """(len(ious), C) fp64 AP table on the device (NaN where a class has no ground truth).  Under torch.distributed
        the records of all ranks are gathered first (global AP on every rank); `gather=False` keeps it local."""
        dev = self.device
        if self._records:
            score = torch.cat([r.score for r in self._records], 0)
            cls = torch.cat([r.cls for r in self._records], 0)
            tp = torch.cat([r.tp for r in self._records], 0)
            counts = torch.cat([r.counts for r in self._records], 0)
        else:
            score = torch.zeros((0, self.max_det), dtype=torch.float32, device=dev)
            cls = torch.zeros((0, self.max_det), dtype=torch.int32, device=dev)
            tp = torch.zeros((0, self.max_det), dtype=torch.uint8, device=dev)
            counts = torch.zeros((0,), dtype=torch.int32, device=dev)
        gt_count = self.gt_count
        if gather:
            score, cls, tp, counts, gt_count = D.gather_records(score, cls, tp, counts, gt_count)
        lib = L.lib()
        n_img = int(counts.shape[0])
        ap = torch.empty((L.NUM_IOU_THRS, self.num_classes), dtype=torch.float64, device=dev)
        ws = H.workspace("ap", dev, lib.kodb_ap_workspace_bytes(n_img, self.max_det, self.num_classes), tag=self.num_classes)
        L.check(lib.kodb_ap_accumulate(score.data_ptr(), cls.data_ptr(), tp.data_ptr(), counts.data_ptr(), n_img,
                                       self.max_det, self.num_classes, gt_count.data_ptr(), ap.data_ptr(),
                                       ws.data_ptr(), ws.numel(), H.stream_ptr(dev)), "kodb_ap_accumulate")
        return ap

    def get_report(self, gather: bool = True) -> dict:
        """Keys consumed at pycoco_map_eval.py:113-125."""
        ap = self.average_precision(gather).cpu()  # the one D2H copy of the epoch
        rep: dict = {}
        per_t = []
        for ti, t in enumerate(self.ious):
            row = ap[ti]
            valid = row[~torch.isnan(row)]
            m = float(valid.mean()) if valid.numel() else -1.0
            per_t.append(m)
            rep[f"mAP_{int(round(t * 100))}"] = m
            if self.report_tag_wise[ti]:
                rep[f"tag_wise_AP_{int(round(t * 100))}"] = {
                    k: float(row[k]) for k in range(self.num_classes) if not bool(torch.isnan(row[k]))}
        rep["avg_mAP"] = sum(per_t) / len(per_t) if per_t else -1.0
        return rep


class PyCOCOMAPEvalCallback(_CallbackBase):
    """Same hooks and logged keys as the reference callback (pycoco_map_eval.py:41-144)."""

    def __init__(self, label_info: dict, device="cuda", **kwargs):
        super().__init__(**kwargs)
        self.label_info = label_info
        self.evaluator = DeviceCocoMapEvaluator(num_classes=len(label_info), ious=[0.3, 0.5, 0.75, 0.9],
                                                report_tag_wise=[False, True, False, False], device=device)

    def on_validation_batch_end(self, trainer, pl_module, outputs: Optional[Any], batch: Any, batch_idx: int,
                                dataloader_idx: int = 0) -> None:
        assert isinstance(outputs, dict)
        if outputs.get("records") is not None:  # fused eval tail already matched this batch
            self.evaluator.add_records(outputs["records"])
            return
        det = outputs["detections"]
        if isinstance(det, NmsResult):
            self.evaluator.add_detections(det.dets, det.counts, outputs["targets"])
        else:
            self.evaluator.add_detection_list(det, outputs["targets"])

    def on_validation_epoch_end(self, trainer, pl_module) -> None:
        report = self.evaluator.get_report()
        results = {"map": report["avg_mAP"], "map30": report["mAP_30"], "map50": report["mAP_50"],
                   "map75": report["mAP_75"], "map90": report["mAP_90"]}
        for k, v in report["tag_wise_AP_50"].items():
            results[f"map50_{self.label_info[k]}"] = v
        try:
            from rich.console import Console
            from rich.table import Table

            table = Table(title=f"MAP [Epoch - {getattr(trainer, 'current_epoch', 0)}]", show_header=False,
                          show_lines=True)
            for k, v in results.items():
                table.add_row(k, str(v))
            Console().print(table)
        except Exception:  # pragma: no cover
            pass
        # the AP is already global (records were gathered across ranks), so no sync_dist averaging
        if hasattr(pl_module, "log_dict"):
            pl_module.log_dict(results, sync_dist=False)
        self.last_results = results
        self.evaluator.reset()
