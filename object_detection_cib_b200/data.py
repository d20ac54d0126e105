"""Target ingest: collate the per-image targets ONCE, in the DataLoader, into the flat
`[boxes | labels | offsets]` blob the kernels read, so that a training / validation step uploads its ground
truth with a single H2D copy of pinned memory and no per-step packing.

* `packed_collate_fn(batch)` – drop-in for `DetectionDataModule.default_collate_fn`
  (kod/lightning/data_module.py:82-87): same `(imgs, targets, image_info)` triple, but `targets` is a
  `HostPackedTargets`.
* `HostPackedTargets` – a `Sequence[DetectionTarget]` (indexing / iteration give per-image views of the blob,
  so reference code that loops over `targets` keeps working) that also carries the flat arrays.  It implements
  `pin_memory()`, which `DataLoader(pin_memory=True)` calls on custom batch objects in the main process.
  `Yolov5Loss`, `nms.eval_tail`, `nms.nms_padded` and the mAP evaluator accept it wherever they accept the
  reference's tuple of targets (`_host.pack_targets`).

dtype rule (yv5.py:92-113): if any image carries fp64 boxes the blob is fp64 and the kernels convert after the
cxcywh + 1/stride arithmetic, exactly as for the tuple form.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .types import DetectionTarget


def _round16(n: int) -> int:
    return (n + 15) // 16 * 16


class HostPackedTargets(Sequence):
    """Flat copy of the targets of one batch: `boxes (T,4)`, `labels (T,) int64`, `offsets (B+1,) int32`,
    all views of ONE uint8 blob `[boxes | labels | offsets]` with 16-byte aligned sections.

    Built on the host by the collate function; `.to(device)` moves the blob with one copy and returns the same
    kind of object, so Lightning's `transfer_batch_to_device` (which calls `.to` on anything that has it instead of
    rebuilding the sequence element by element) carries it to the GPU intact."""

    def __init__(self, blob: torch.Tensor, counts: Sequence[int], f64: bool):
        self.blob = blob
        self.counts = [int(c) for c in counts]
        self.f64 = bool(f64)
        T, B = sum(self.counts), len(self.counts)
        bsz = 8 if f64 else 4
        self.nb_boxes = _round16(T * 4 * bsz)
        self.nb_labels = _round16(T * 8)
        self.nb_offsets = _round16((B + 1) * 4)
        if blob.numel() < max(self.nb_boxes + self.nb_labels + self.nb_offsets, 16):
            raise ValueError("blob too small for the given counts")
        bdt = torch.float64 if f64 else torch.float32
        self.boxes = blob[:T * 4 * bsz].view(bdt).view(T, 4)
        self.labels = blob[self.nb_boxes:self.nb_boxes + T * 8].view(torch.long)
        self.offsets = blob[self.nb_boxes + self.nb_labels:self.nb_boxes + self.nb_labels + (B + 1) * 4].view(torch.int32)

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_targets(cls, targets: Sequence, pin: bool = False) -> "HostPackedTargets":
        """tuple of B `DetectionTarget(boxes (n_i,4) xyxy px, labels (n_i,))` (possibly empty `(0,)`-shaped arrays,
        sample_reader.py:133) -> one blob.  Lengths come from `.shape`."""
        counts = [int(t[0].shape[0]) if t[0].dim() > 0 and t[0].numel() else 0 for t in targets]
        boxes_l = [t[0].reshape(-1, 4) for t, n in zip(targets, counts) if n]
        labels_l = [t[1].reshape(-1) for t, n in zip(targets, counts) if n]
        f64 = any(b.dtype == torch.float64 for b in boxes_l)
        T, B = sum(counts), len(counts)
        bsz = 8 if f64 else 4
        nbytes = max(_round16(T * 4 * bsz) + _round16(T * 8) + _round16((B + 1) * 4), 16)
        blob = torch.zeros(nbytes, dtype=torch.uint8)
        if pin:
            blob = blob.pin_memory()
        out = cls(blob, counts, f64)
        if T:
            bdt = torch.float64 if f64 else torch.float32
            torch.cat([b.detach().to("cpu", bdt) for b in boxes_l], 0, out=out.boxes)
            torch.cat([l.detach().to("cpu", torch.long) for l in labels_l], 0, out=out.labels)
        offs = np.zeros(B + 1, np.int32)
        np.cumsum(counts, out=offs[1:])
        out.offsets.copy_(torch.from_numpy(offs))
        return out

    def pin_memory(self) -> "HostPackedTargets":
        """DataLoader(pin_memory=True) protocol for custom batch types."""
        if self.blob.is_cuda or self.blob.is_pinned():
            return self
        return HostPackedTargets(self.blob.pin_memory(), self.counts, self.f64)

    def to(self, device=None, *args, non_blocking: bool = True, **kwargs) -> "HostPackedTargets":
        """Move the blob (ONE copy) and keep the packed form.  Having `.to` makes the object a transferable leaf for
        `lightning.pytorch.utilities.move_data_to_device` / `apply_to_collection`, which would otherwise try to
        rebuild this Sequence as `type(self)([...])`.  dtype arguments are ignored (boxes keep their precision)."""
        if device is None or isinstance(device, torch.dtype):
            return self
        device = torch.device(device)
        if device == self.blob.device:
            return self
        return HostPackedTargets(self.blob.to(device, non_blocking=non_blocking), self.counts, self.f64)

    @property
    def device(self) -> torch.device:
        return self.blob.device

    # ---- Sequence[DetectionTarget] ----------------------------------------------------------------------
    def __len__(self) -> int:
        return len(self.counts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        lo = sum(self.counts[:i])  # host-side lengths: no device sync when the blob lives on the GPU
        hi = lo + self.counts[i]
        return DetectionTarget(self.boxes[lo:hi], self.labels[lo:hi])

    @property
    def num_targets(self) -> int:
        return sum(self.counts)


def packed_collate_fn(batch):
    """Drop-in for `DetectionDataModule.default_collate_fn` (kod/lightning/data_module.py:82-87)."""
    imgs, target, image_info = zip(*batch)
    return torch.stack(imgs, dim=0), HostPackedTargets.from_targets(target), image_info
