"""Probe of the A-operand descriptor of kodb_head_forward (csrc/head.cu): each combination in its own process
(KODB_HD_DBG = "tma swizzle, UMMA layout type, LBO, SBO, k advance"), error against a strict-fp32 matmul."""
import os
import subprocess
import sys

CHILD = r"""
import sys, torch
sys.path.insert(0, ".")
from object_detection_cib_b200.heads import Yolov5Head
dev = "cuda:0"
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
for (B, K, Hh, Ww, C) in ((1, 32, 16, 8, 10), (2, 128, 20, 20, 10), (2, 256, 12, 12, 80)):
    h = Yolov5Head(K, 3, C, 8).to(dev)
    x = torch.randn(B, K, Hh, Ww, device=dev)
    with torch.no_grad():
        r = h(x)
        errs = []
        for t, m in zip(r, (h.box_head, h.obj_head, h.cls_head)):
            w = m.conv.weight.reshape(-1, K)
            want = torch.einsum("nk,bkp->bnp", w, x.reshape(B, K, -1)) + m.conv.bias[None, :, None]
            got = t.permute(0, 1, 4, 2, 3).reshape(B, -1, Hh * Ww)
            errs.append(float((got - want).abs().max()))
    print("  shape", (B, K, Hh, Ww, C), "max abs err box/obj/cls:", ["%.3g" % e for e in errs], "nonzero:", bool(r[0].abs().sum() > 0))
"""

combos = sys.argv[1:] or ["", "4,1,4096,512,1024", "4,1,4096,1024,1024", "4,1,512,4096,1024", "3,2,4096,1024,1024"]
for c in combos:
    env = dict(os.environ)
    if c:
        env["KODB_HD_DBG"] = c
    print("combo", c or "(default)", flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=120)
    print(r.stdout, r.stderr[-600:] if r.returncode else "", flush=True)
