"""Multi-GPU plumbing for the eval tail.

The detection tail shards by image: every rank runs the kernels on its own images and there is no
data-path collective.  The only exchange is at validation epoch end: the per-detection records
(score, class, TP bits) and the per-class ground-truth counts are gathered so that every rank
accumulates the *global* AP.  (The reference instead logs per-rank mAPs with `sync_dist=True`,
i.e. their mean – pycoco_map_eval.py:139-142.)  Works with NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """all_gather of tensors whose first dimension differs per rank; result is rank-major."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.long, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    if mx == 0:
        return t.new_zeros((0,) + tuple(t.shape[1:]))
    pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
    pad[: t.shape[0]] = t
    # uint8 / int32 / float32 all travel as-is on both NCCL and gloo
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


def gather_records(score: torch.Tensor, cls: torch.Tensor, tp: torch.Tensor, counts: torch.Tensor,
                   gt_count: torch.Tensor, group=None):
    """Per-rank (num_images_r, max_det) record planes + (C,) GT counts -> global planes on every rank."""
    if not is_distributed():
        return score, cls, tp, counts, gt_count
    score = gather_rows(score.contiguous(), group)
    cls = gather_rows(cls.contiguous(), group)
    tp = gather_rows(tp.contiguous(), group)
    counts = gather_rows(counts.contiguous(), group)
    total = gt_count.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return score, cls, tp, counts, total


def shard_range(num_items: int, rank: int, world: int) -> range:
    """Contiguous image shard of a batch/dataset for one rank (weak scaling keeps per-rank size fixed)."""
    per = (num_items + world - 1) // world
    return range(min(rank * per, num_items), min((rank + 1) * per, num_items))
