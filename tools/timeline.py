#!/usr/bin/env python
"""Timeline of the pipelined step (debug): eager launches on the bench's stream layout (one branch per input set, train tail on
a side stream) with a CUDA event around every kernel; prints, for a few steps in the middle, when each kernel became eligible
and when it finished, relative to the start.  usage: timeline.py [steps] [kinds: subset of c,s,t]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from object_detection_cib_b200 import _lib as L
from object_detection_cib_b200 import synthetic as syn
from object_detection_cib_b200.iou import IoUCalculator, IoUType
from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams
from object_detection_cib_b200.map_eval import DeviceCocoMapEvaluator
from object_detection_cib_b200.types import default_anchor_info

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kinds = sys.argv[2] if len(sys.argv) > 2 else "cst"
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
info = default_anchor_info()
env = dict(dev=dev, lib=L.lib(), rank=0, world=1, info=info,
           make_loss=lambda Cn: Yolov5Loss(Yolov5LabelAssigner(info), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou, 1e-7),
                                           syn.readme_class_weights() if Cn == 10 else None),
           make_evaluator=lambda Cn: DeviceCocoMapEvaluator(Cn, device=dev))
wl = bench.TailWorkload(env, 64, 640, 10)
ns = len(wl.sets)
for i in range(ns):
    wl.step_serial(i)
torch.cuda.synchronize()
branches = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(ns)]
sides = [torch.cuda.Stream(device=dev) for _ in range(ns)]
E = lambda: torch.cuda.Event(enable_timing=True)
base = E()
rec = []  # (step, kind, ev0, ev1)
main = torch.cuda.current_stream(dev)
torch.cuda._sleep(int(0.08 * 1.9e9))  # the host queues everything behind this
base.record(main)
for s in branches + sides:
    s.wait_stream(main)
for i in range(steps if "P" not in kinds else 0):
    st = wl.sets[i % ns]
    br, sd = branches[i % ns], sides[i % ns]
    with torch.cuda.stream(br):
        if "c" in kinds:
            e0, e1 = E(), E()
            e0.record(br); wl.eval_call(st, 1); e1.record(br)
            rec.append((i, "cand", e0, e1))
        fork = E()
        fork.record(br)
        if "s" in kinds:
            e0, e1 = E(), E()
            e0.record(br); wl.eval_call(st, 2); e1.record(br)
            rec.append((i, "select", e0, e1))
    if "t" in kinds:
        sd.wait_event(fork)
        with torch.cuda.stream(sd):
            e0, e1 = E(), E()
            e0.record(sd); wl.train_call(st); e1.record(sd)
            rec.append((i, "train", e0, e1))
        br.wait_stream(sd)
if "P" in kinds:
    # phased schedule: the candidate kernels of `ns` steps together, then their select/NMS kernels and train tails together
    rec = []
    for g0 in range(0, steps, ns):
        for i in range(g0, min(g0 + ns, steps)):
            with torch.cuda.stream(branches[i % ns]):
                wl.eval_call(wl.sets[i % ns], 1)
        join = [E() for _ in range(ns)]
        for k in range(ns):
            join[k].record(branches[k])
        for s in branches + sides:
            for k in range(ns):
                s.wait_event(join[k])
        for i in range(g0, min(g0 + ns, steps)):
            with torch.cuda.stream(branches[i % ns]):
                wl.eval_call(wl.sets[i % ns], 2)
            with torch.cuda.stream(sides[i % ns]):
                wl.train_call(wl.sets[i % ns])
        join2 = [E() for _ in range(2 * ns)]
        for k, s in enumerate(branches + sides):
            join2[k].record(s)
        for s in branches + sides:
            for k in range(2 * ns):
                s.wait_event(join2[k])
for s in branches + sides:
    main.wait_stream(s)
end = E()
end.record(main)
torch.cuda.synchronize()
print(f"kinds={kinds}: {steps} steps in {base.elapsed_time(end) * 1e3:.0f} us -> {base.elapsed_time(end) * 1e3 / steps:.1f} us/step (eager, event-bracketed)")
lo, hi = steps // 2, steps // 2 + 7
for i, kind, e0, e1 in rec:
    if lo <= i < hi:
        t0, t1 = base.elapsed_time(e0) * 1e3, base.elapsed_time(e1) * 1e3
        print(f"step {i:3d} set {i % ns} {kind:6s} eligible {t0:8.1f} done {t1:8.1f}  ({t1 - t0:6.1f} us)")
