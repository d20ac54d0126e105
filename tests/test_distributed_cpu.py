"""CPU, world_size 2, gloo: the only exchange of the path – gathering per-detection records and GT counts at
validation epoch end – plus the contiguous image sharding helper."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from object_detection_cib_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_img = 3 + 2 * rank  # ragged: ranks hold different numbers of images
        g = torch.Generator().manual_seed(100 + rank)
        score = torch.rand(n_img, 300, generator=g)
        cls = torch.randint(0, 10, (n_img, 300), generator=g, dtype=torch.int32)
        tp = torch.randint(0, 16, (n_img, 300), generator=g, dtype=torch.uint8)
        counts = torch.randint(0, 301, (n_img,), generator=g, dtype=torch.int32)
        gt = torch.arange(10, dtype=torch.long) * (rank + 1)
        s2, c2, t2, n2, g2 = D.gather_records(score, cls, tp, counts, gt)
        torch.save({"score": s2, "cls": c2, "tp": t2, "counts": n2, "gt": g2,
                    "mine": (score, cls, tp, counts, gt)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_gather_records_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    want = [torch.cat([o["mine"][i] for o in outs], 0) for i in range(4)]
    want_gt = sum(o["mine"][4] for o in outs)
    for o in outs:  # every rank ends up with the same global, rank-major view
        assert torch.equal(o["score"], want[0]) and torch.equal(o["cls"], want[1])
        assert torch.equal(o["tp"], want[2]) and torch.equal(o["counts"], want[3])
        assert torch.equal(o["gt"], want_gt)
        assert o["score"].shape[0] == 3 + 5


def test_not_distributed_is_identity():
    a = torch.zeros(2, 300)
    out = D.gather_records(a, a.int(), a.to(torch.uint8), torch.zeros(2, dtype=torch.int32), torch.zeros(10, dtype=torch.long))
    assert out[0] is a


def test_shard_range_covers_everything():
    for n, w in ((64, 8), (10, 4), (3, 8), (128, 2)):
        seen = []
        for r in range(w):
            seen += list(D.shard_range(n, r, w))
        assert seen == list(range(n))
