#!/usr/bin/env python
"""Fused Detect head (kodb_head_forward) against the reference's three cuDNN 1x1 convolutions per level, same box.
usage: bench_head.py [B S C]"""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from object_detection_cib_b200 import nms as N
from object_detection_cib_b200.heads import Yolov5Head, fused_heads_detections, fused_heads_forward
from object_detection_cib_b200.types import FeatureShape, default_anchor_info

B, S, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 640, 10)
dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = ((128, 8), (256, 16), (512, 32))
heads = [Yolov5Head(K, 3, C, s).to(dev) for K, s in cfg]
NSETS = 3
xsets = [[torch.randn(B, K, S // s, S // s, device=dev) for K, s in cfg] for _ in range(NSETS)]


def ref(xs):
    out = []
    for h, x in zip(heads, xs):
        out.append([F.conv2d(x, m.conv.weight, m.conv.bias) for m in (h.box_head, h.obj_head, h.cls_head)])
    return out


def timed(fn, n=30):
    """device time per call: n calls (inputs rotating over NSETS sets, > L2 together) captured in ONE CUDA graph"""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(5):
            fn(xsets[i % NSETS])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                fn(xsets[i % NSETS])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    t_ours = timed(lambda xs: fused_heads_forward(heads, xs))
    torch.backends.cudnn.allow_tf32 = True
    t_ref_tf32 = timed(ref)
    torch.backends.cudnn.allow_tf32 = False
    t_ref_fp32 = timed(ref)
    got = fused_heads_forward(heads, xsets[0])
    want = ref(xsets[0])
    err = max(float((g.permute(0, 1, 4, 2, 3).reshape(w.shape) - w).abs().max()) for gl, wl in zip(got, want) for g, w in zip(gl, wl))
with torch.no_grad():
    info, shape = default_anchor_info(), FeatureShape(S, S)
    t_det = timed(lambda xs: fused_heads_detections(heads, xs, shape, info))
    t_two = timed(lambda xs: N.get_detections(shape, fused_heads_forward(heads, xs), info))
n45 = 3 * (5 + C)
bytes_in = sum(4 * B * K * (S // s) ** 2 for K, s in cfg)
bytes_out = sum(4 * B * n45 * (S // s) ** 2 for K, s in cfg)
print(f"head B={B} S={S} C={C}: fused {t_ours:.1f} us ({(bytes_in + bytes_out) / t_ours / 1e3:.0f} GB/s of {(bytes_in + bytes_out) / 1e6:.0f} MB, "
      f"frac of 6551 = {(bytes_in + bytes_out) / t_ours / 1e3 / 6551:.3f}); 3 cuDNN convs/level: tf32 {t_ref_tf32:.1f} us, fp32 {t_ref_fp32:.1f} us; "
      f"speed-up {t_ref_tf32 / t_ours:.2f}x / {t_ref_fp32 / t_ours:.2f}x; max |err| vs fp32 {err:.2e}")
print(f"  features -> detections (decode epilogue): {t_det:.1f} us ({(bytes_in + bytes_out) / t_det / 1e3:.0f} GB/s, frac {(bytes_in + bytes_out) / t_det / 1e3 / 6551:.3f}); "
      f"planar head + decode kernel: {t_two:.1f} us")
