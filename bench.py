#!/usr/bin/env python
"""Benchmark of the YOLOv5s detection tail (BASELINE.json metric: head+loss+NMS images/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

One "step" = one pass of the hot path over one batch of `--batch` images per GPU (default: config[1] of
BASELINE.json, 64 images/GPU at 640x640, 10 classes, class-weighted loss):
  train tail : target assignment + CIoU/objectness/class-weighted BCE loss, forward AND backward
               (kodb_train_tail, one kernel)
  eval tail  : sigmoid/grid/anchor decode + conf 0.001 filter + class-offset NMS (IoU 0.6, 300/img)
               + per-class COCO matching (kodb_eval_tail, two kernels)
`value` is device-resident throughput (CUDA events, max over ranks); `e2e` goes through the public
Python drop-ins with host buffers (H2D of the head maps and targets, D2H of losses/detections/records
inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "yolov5s_tail_images_per_s"
UNIT = "images/s"
NSETS = int(os.environ.get("KODB_BENCH_NSETS", "3"))  # rotating input sets: 3 x (96.8 MB head + 96.8 MB grads) >> 126 MB L2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="images per GPU")
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fork", action="store_true", help="serialise the train and eval tails inside the graph")
    ap.add_argument("--fork", choices=("late", "early"), default="late",
                    help="late: candidates -> (select/NMS || train tail); early: eval tail || train tail from the start")
    ap.add_argument("--pipeline", type=int, default=1, choices=(0, 1),
                    help="1: replay the graph of input set s on its own stream (consecutive steps overlap); 0: one stream")
    ap.add_argument("--train-delay-cycles", type=int, default=0,
                    help="--fork late: spin this many SM cycles on the train stream before the train tail is launched")
    ap.add_argument("--cpu-sample", type=int, default=8, help="images in the cpu_baseline sample")
    return ap.parse_args()


def workload_name(a):
    return (f"yolov5s detection tail, {a.batch} img/GPU {a.img}x{a.img}, {a.classes} classes: train tail "
            f"(assign + CIoU/obj/cls loss fwd+bwd, class weights) + eval tail (decode, conf 0.001, NMS 0.6, "
            f"mAP matching); synthetic Zipf boxes")


def algorithmic_bytes(a):
    """per-image algorithmic bytes of the two tails."""
    ac = sum(3 * (a.img // s) ** 2 for s in (8, 16, 32))
    train = 4 * ac * (1 + 5 + a.classes)
    evalb = 4 * ac * (5 + a.classes) + 300 * (6 * 4 + 8)
    return ac, train, evalb


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the loaded regions
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if not (t0 - 0.05 <= ts <= t1 + 0.1):
                continue
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, v in zip(self.NAMES, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from object_detection_cib_b200 import _host as H
    from object_detection_cib_b200 import _lib as L
    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200 import synthetic as syn
    from object_detection_cib_b200.iou import IoUCalculator, IoUType
    from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
    from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams, launch_train_tail
    from object_detection_cib_b200.map_eval import DeviceCocoMapEvaluator
    from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    B, S, Cn = a.batch, a.img, a.classes
    shape, info = FeatureShape(S, S), default_anchor_info()
    weights = syn.readme_class_weights() if Cn == 10 else None
    loss_mod = Yolov5Loss(Yolov5LabelAssigner(info), Yolov5LossParams.get_default(), IoUCalculator(IoUType.ciou, 1e-7),
                          weights)
    evaluator = DeviceCocoMapEvaluator(Cn, device=dev)

    # ---- NSETS rotating input sets, resident in HBM ---------------------------------------------------
    sets = []
    for s in range(NSETS):
        seed = 2023 + 17 * s + 101 * rank
        heads = syn.make_head(B, S, Cn, seed=seed, device=dev)
        targets = syn.make_targets(B, S, Cn, seed=seed, n_min=0, n_max=10)
        syn.inject_ground_truth(heads, targets, S, fraction=0.3, seed=seed)
        net = [DetectionHeadResult(*lv) for lv in heads]
        plan = loss_mod.plan(shape, net, targets)
        # one train workspace per input set: steps of different sets may be in flight at the same time (--pipeline)
        plan = plan._replace(workspace=H.workspace(f"train{s}", dev, plan.workspace.numel()))
        grads = [H.alloc_planar_like(t) for t in plan.heads.tensors]
        losses = torch.zeros(3, dtype=torch.float32, device=dev)
        nass = torch.zeros(3, dtype=torch.int32, device=dev)
        levels = H.level_cfgs(info, S, S)
        rows = sum(lv.num_anchors * lv.H * lv.W for lv in levels)
        cfg = N._nms_cfg(Cn, 0.001, 0.6, None, N.MAX_DET, N.MAX_NMS)
        res, out = N._alloc_out(B, N.MAX_DET, dev)
        match = N._alloc_match(B, N.MAX_DET, evaluator.gt_count, res.counts, dev, N.IOU_THRESHOLDS)
        ws = H.workspace(f"eval{s}", dev, lib.kodb_eval_workspace_bytes(B, rows, Cn))
        sets.append(dict(heads=heads, targets=targets, net=net, plan=plan, grads=grads, losses=losses, nass=nass,
                         levels=levels, cfg=cfg, res=res, out=out, match=match, ws=ws, gstruct=H.grad_structs(grads)))
    hint = (float(B),) * 3
    up = (C.c_float * 3)(*hint)
    errs = []

    def train_call(st):
        p = st["plan"]
        rc = lib.kodb_train_tail(p.heads.structs, p.levels, C.byref(p.packed.struct), C.byref(p.cfg),
                                 p.pos_weight.data_ptr() if p.pos_weight is not None else None, up, None, 0,
                                 st["gstruct"], st["losses"].data_ptr(), st["nass"].data_ptr(), p.workspace.data_ptr(),
                                 p.workspace.numel(), H.stream_ptr(dev))
        if rc:
            errs.append(("train", rc))

    def eval_call(st, stage_mask=0):
        p = st["plan"]
        st["cfg"].stage_mask = stage_mask
        rc = lib.kodb_eval_tail(p.heads.structs, st["levels"], B, C.byref(st["cfg"]), C.byref(st["out"]),
                                C.byref(p.packed.struct), C.byref(st["match"].struct), st["ws"].data_ptr(),
                                st["ws"].numel(), H.stream_ptr(dev))
        st["cfg"].stage_mask = 0
        if rc:
            errs.append(("eval", rc))

    def step_eager(i):
        st = sets[i % NSETS]
        train_call(st)
        eval_call(st)

    # The two tails of a step are independent (both only read the head maps): inside the CUDA graph they are
    # forked onto two streams, so the 64-CTA NMS kernel (one CTA per image) overlaps the streaming kernels.
    side = torch.cuda.Stream(device=dev)
    fork_ev = torch.cuda.Event()

    def step_forked(i):
        # candidate kernel (all SMs, HBM-bound) first; then the 64-CTA select/NMS kernel (one CTA per image, latency
        # bound, one full SM each) and the train tail (HBM-bound stream) share the GPU
        st = sets[i % NSETS]
        main = torch.cuda.current_stream(dev)
        if a.fork == "early":
            side.wait_stream(main)
            with torch.cuda.stream(side):
                eval_call(st)
            train_call(st)
            main.wait_stream(side)
            return
        eval_call(st, 1)
        fork_ev.record(main)
        eval_call(st, 2)
        side.wait_event(fork_ev)
        with torch.cuda.stream(side):
            if a.train_delay_cycles > 0:
                torch.cuda._sleep(a.train_delay_cycles)  # let the select kernel's one-per-SM CTAs get their SMs first
            train_call(st)
        main.wait_stream(side)

    LAUNCHES_PER_STEP = 3  # train_tail_kernel, cand_floor_kernel, select_nms_kernel

    # ---- CUDA graphs: one per input set -----------------------------------------------------------------
    for i in range(NSETS):  # first launches set per-device function attributes (not capturable)
        step_eager(i)
    torch.cuda.synchronize(dev)
    graphs = None
    # the capture stream carries the select/NMS kernel: high priority, so that its 64 one-per-SM CTAs are placed
    # before the train tail's many small CTAs when both become ready behind the candidate kernel
    cap_stream = torch.cuda.Stream(device=dev, priority=-1)
    if not a.no_graph:
        try:
            graphs = []
            for i in range(NSETS):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream):
                    (step_eager if a.no_fork else step_forked)(i)
                graphs.append(g)
        except Exception as e:  # pragma: no cover
            sys.stderr.write(f"[bench] CUDA graph capture failed ({e}); timing eager launches\n")
            graphs = None
            torch.cuda.synchronize(dev)

    # Steps are independent (each input set has its own buffers, outputs and workspaces): with --pipeline the graph
    # of set s is always replayed on replay stream s, so up to NSETS consecutive steps are in flight and the
    # latency-bound select/NMS kernel of one step overlaps the streaming kernels of the next.
    rstreams = [torch.cuda.Stream(device=dev) for _ in range(NSETS)] if (a.pipeline and graphs is not None) else None

    def step(i):
        if graphs is not None:
            if rstreams is not None:
                with torch.cuda.stream(rstreams[i % NSETS]):
                    graphs[i % NSETS].replay()
            else:
                graphs[i % NSETS].replay()
        else:
            step_eager(i)

    def fork_replay_streams():
        if rstreams is not None:
            cur = torch.cuda.current_stream(dev)
            for r in rstreams:
                r.wait_stream(cur)

    def join_replay_streams():
        if rstreams is not None:
            cur = torch.cuda.current_stream(dev)
            for r in rstreams:
                cur.wait_stream(r)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)

    # ---- warm-up + the timed region: EXACTLY K steps ------------------------------------------------------
    fork_replay_streams()
    for i in range(max(a.warmup, 3)):
        step(i)
    join_replay_streams()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.time()
    e0.record()
    fork_replay_streams()  # the replay streams start behind e0 ...
    for i in range(a.steps):
        step(i)
    join_replay_streams()  # ... and e1 is recorded behind the last kernel of every one of them
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    ms_per_step = ms_total / a.steps
    value = world * B * a.steps / (ms_total / 1e3)

    # ---- instrumented pass: per-kernel device time ------------------------------------------------------
    # Eager launches with CUDA events between the kernels.  The host needs longer to issue a launch through
    # ctypes than these kernels run, so the whole pass is enqueued behind a ~40 ms spin kernel: by the time
    # the GPU reaches it everything is queued and the events see back-to-back device time only.
    n_inst = min(max(a.steps, 20), 60)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_inst)]
    for i in range(3):
        step_eager(i)
    torch.cuda.synchronize(dev)
    torch.cuda._sleep(int(0.04 * 1.9e9))
    for i in range(n_inst):
        st = sets[i % NSETS]
        evs[i][0].record()
        train_call(st)
        evs[i][1].record()
        eval_call(st, 1)
        evs[i][2].record()
        eval_call(st, 2)
        evs[i][3].record()
    torch.cuda.synchronize(dev)
    t_train_ev = statistics.median(e[0].elapsed_time(e[1]) for e in evs)
    t_cand_ev = statistics.median(e[1].elapsed_time(e[2]) for e in evs)
    t_sel_ev = statistics.median(e[2].elapsed_time(e[3]) for e in evs)
    # An event between two launches costs 2-4 us of idle GPU per interval.  The per-kernel figures used for the
    # roofline are therefore taken from back-to-back launches of the same work between ONE pair of events (behind the
    # spin kernel again): n_inst train tails, then n_inst eval tails (candidate + select kernels; their pair time is
    # split in the ratio of the event-separated measurements above).
    bb = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda._sleep(int(0.04 * 1.9e9))
    bb[0].record()
    for i in range(n_inst):
        train_call(sets[i % NSETS])
    bb[1].record()
    for i in range(n_inst):
        eval_call(sets[i % NSETS])
    bb[2].record()
    torch.cuda.synchronize(dev)
    t_load1 = time.time()
    t_train = bb[0].elapsed_time(bb[1]) / n_inst
    t_pair = bb[1].elapsed_time(bb[2]) / n_inst
    t_cand = t_pair * t_cand_ev / (t_cand_ev + t_sel_ev)
    t_sel = t_pair - t_cand

    # ---- end-to-end through the public drop-in API with HOST buffers -----------------------------------
    e2e = None
    if not a.no_e2e:
        st = sets[0]
        host_heads = [[t.permute(0, 1, 4, 2, 3).contiguous().cpu().pin_memory() for t in lv] for lv in st["heads"]]
        host_out = dict(
            losses=torch.empty(3, dtype=torch.float32).pin_memory(),
            dets=torch.empty((B, N.MAX_DET, 6), dtype=torch.float32).pin_memory(),
            counts=torch.empty((B,), dtype=torch.int32).pin_memory(),
            tp=torch.empty((B, N.MAX_DET), dtype=torch.uint8).pin_memory())
        h2d = sum(t.numel() * 4 for lv in host_heads for t in lv) + sum(b.numel() * 4 + l.numel() * 8 for b, l in st["targets"])
        d2h = sum(t.numel() * t.element_size() for t in host_out.values())
        ev_e2e = DeviceCocoMapEvaluator(Cn, device=dev)

        # Double-buffered: the H2D copy of step i+1 (copy stream) overlaps the kernels and the D2H of step i; the host
        # waits for the results of step i-1 before it issues step i+1's copies, and for the last one before the clock
        # stops -- every step's inputs cross PCIe and every step's results reach the host inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        dev_in = [[[torch.empty_like(t, device=dev) for t in lv] for lv in host_heads] for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def e2e_upload(k):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[k % 2])  # the buffers' previous step has been consumed
                for lv_d, lv_h in zip(dev_in[k % 2], host_heads):
                    for td, th in zip(lv_d, lv_h):
                        td.copy_(th, non_blocking=True)
                ready[k % 2].record(copy_stream)

        def e2e_compute(k):
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready[k % 2])
            net = [DetectionHeadResult(*[t.permute(0, 1, 3, 4, 2).detach().requires_grad_(True) for t in lv]) for lv in dev_in[k % 2]]
            r = loss_mod(shape, net, st["targets"])  # packs + uploads the targets
            total = B * (r.localization + r.classification + r.objectness)  # exp.py:124-130
            total.backward()
            res, rec = N.eval_tail(shape, net, info, 0.001, 0.6, targets=st["targets"], gt_count=ev_e2e.gt_count)
            host_out["losses"].copy_(torch.stack((r.localization, r.objectness, r.classification)).detach(), non_blocking=True)
            host_out["dets"].copy_(res.dets, non_blocking=True)
            host_out["counts"].copy_(res.counts, non_blocking=True)
            host_out["tp"].copy_(rec.tp, non_blocking=True)
            done[k % 2].record(cur)

        def e2e_run(n):
            e2e_upload(0)
            for k in range(n):
                if k + 1 < n:
                    e2e_upload(k + 1)
                e2e_compute(k)
                if k >= 1:
                    done[(k - 1) % 2].synchronize()  # step k-1's results are on the host
            torch.cuda.synchronize(dev)

        n_e2e = max(3, min(a.steps, 30))
        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(n_e2e)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * n_e2e / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "api": "Yolov5Loss.forward+backward, nms.eval_tail (pinned host head maps -> device -> pinned host results; H2D of step i+1 overlaps step i)"}

    # ---- epoch-end exchange (the only collective of the path): gather records + global AP ---------------
    evaluator.reset()  # the timed loops accumulated GT counts for thousands of steps; start a clean "epoch"
    for st in sets:
        st["match"] = N._alloc_match(B, N.MAX_DET, evaluator.gt_count, st["res"].counts, dev, N.IOU_THRESHOLDS)
        eval_call(st)
        evaluator.add_records(st["match"].records)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    report = evaluator.get_report()
    epoch_end_ms = (time.perf_counter() - t0) * 1e3

    # diagnostic: share of the candidate kernel's warp tiles that listed ALL their candidates in the last eval call (tiles
    # that ran before the image's score floor was published, plus the sample tiles); eval.cu carve_ws layout
    full_tiles = None
    try:
        tiles = sum(lv.num_anchors * ((lv.H * lv.W + 127) // 128) for lv in sets[0]["levels"])
        groups = (Cn + 15) // 16
        al = lambda v: (v + 255) // 256 * 256
        off = al(B * 4) + al(B * 512 * 4) + al(B * 4)
        flags = sets[0]["ws"][off:off + B * tiles * groups * 4].view(torch.int32)
        full_tiles = float((flags == 511).float().mean().item())  # chunk floor HB - 1: listed with no floor
    except Exception:
        full_tiles = None

    clocks = sampler.stop(t_load0, t_load1) if rank == 0 else None
    if errs:
        raise RuntimeError(f"C-ABI calls failed: {errs[:3]}")
    if rank == 0:
        ac, b_train, b_eval = algorithmic_bytes(a)
        peak, peak_src = measured_peak_gbs()
        # per-kernel split of SURVEY §8(d)'s per-image figures: the candidate kernel reads the obj +
        # class planes, the select/NMS kernel owns the box planes and the outputs
        b_cand = 4 * ac * (1 + Cn)
        b_sel = b_eval - b_cand
        kernels = {
            "train_tail_kernel": (t_train, b_train * B),
            "cand_floor_kernel": (t_cand, b_cand * B),
            "select_nms_kernel": (t_sel, b_sel * B),
        }
        # The roofline object describes the dominant HBM-bound kernel: of the two streaming kernels (they move 87% of
        # the step's algorithmic bytes) the one with the longer launch.  select_nms_kernel is one CTA per image -- a
        # latency chain on 64 of 148 SMs, not a bandwidth kernel; its time and (box planes + outputs) bytes are listed
        # in roofline_stages all the same, and roofline_stages.step is the whole path against the HBM peak.
        dom = max(("train_tail_kernel", "cand_floor_kernel"), key=lambda k: kernels[k][0])
        t_dom, bytes_dom = kernels[dom]
        ach = bytes_dom / (t_dom * 1e-3) / 1e9
        traffic = None
        try:  # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("images_per_launch") == B:
                traffic = tr["kernels"].get(dom)
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "images_per_gpu": B, "global_batch": B * world, "img": S,
                       "classes": Cn, "parallelism": f"dp{world} (shard by image, no data-path collective)",
                       "l2": f"inputs rotate over {NSETS} sets ({NSETS}x{(b_eval * B) / 1e6:.0f} MB head + grads) > 126 MB L2",
                       "launch": (("cuda-graph replay: candidate kernel, then select/NMS kernel || train tail on two captured streams"
                                   if a.fork == "late" else "cuda-graph replay, train tail || eval tail on two captured streams")
                                  if not a.no_fork else "cuda-graph replay, serial") + (
                                      f"; consecutive steps pipelined over {NSETS} replay streams (one per input set)" if a.pipeline else "")
                       if graphs is not None else "eager async launches"},
            "gpu_launches": LAUNCHES_PER_STEP * a.steps,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_dom, "kernel_ms": t_dom},
            "stages_ms": {"train_tail_kernel": t_train, "cand_floor_kernel": t_cand, "select_nms_kernel": t_sel,
                          "note": "device time per launch from back-to-back eager launches between one pair of CUDA events, queued "
                                  "behind a spin kernel (train tail alone; eval tail as candidate+select pairs, split in the ratio "
                                  "of the event-separated times)",
                          "event_separated_ms": {"train_tail_kernel": t_train_ev, "cand_floor_kernel": t_cand_ev,
                                                 "select_nms_kernel": t_sel_ev}},
            "roofline_stages": {
                "train_tail": {"achieved_gbs": b_train * B / (t_train * 1e-3) / 1e9, "frac": b_train * B / (t_train * 1e-3) / 1e9 / peak},
                "cand_floor_kernel": {"achieved_gbs": b_cand * B / (t_cand * 1e-3) / 1e9, "frac": b_cand * B / (t_cand * 1e-3) / 1e9 / peak},
                "select_nms_kernel": {"bound": "latency (one CTA per image)", "achieved_gbs": b_sel * B / (t_sel * 1e-3) / 1e9,
                                      "frac": b_sel * B / (t_sel * 1e-3) / 1e9 / peak},
                "eval_tail": {"achieved_gbs": b_eval * B / ((t_cand + t_sel) * 1e-3) / 1e9,
                              "frac": b_eval * B / ((t_cand + t_sel) * 1e-3) / 1e9 / peak},
                "step": {"achieved_gbs": (b_train + b_eval) * B / (ms_per_step * 1e-3) / 1e9,
                         "frac": (b_train + b_eval) * B / (ms_per_step * 1e-3) / 1e9 / peak}},
            "e2e": e2e, "clocks": clocks, "epoch_end_ms": epoch_end_ms, "tiles_listed_in_full": full_tiles,
            "map_report": {k: report[k] for k in ("avg_mAP", "mAP_50")},
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, a.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference path (the ONLY place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------------
def _cpu_inputs(a, n_img, seed=2023):
    import torch

    from object_detection_cib_b200 import synthetic as syn

    heads = syn.make_head(n_img, a.img, a.classes, seed=seed)
    targets = syn.make_targets(n_img, a.img, a.classes, seed=seed, n_min=0, n_max=10)
    syn.inject_ground_truth(heads, targets, a.img, fraction=0.3, seed=seed)
    return heads, targets


def _cpu_step(a, heads, targets, nms_backend):
    """One pass of the reference algorithm over `heads`: train tail fwd+bwd, decode, NMS, matching."""
    import numpy as np

    from object_detection_cib_b200 import synthetic as syn
    from oracle import assign as oa
    from oracle import coco_map as om
    from oracle import decode as od
    from oracle import loss as ol
    from oracle import nms as on

    levels = list(zip(syn.STRIDES, syn.ANCHORS_PX))
    n = heads[0][0].shape[0]
    asg = oa.assign([(b.numpy(), l.numpy()) for b, l in targets], levels, a.img, a.img)
    w = syn.readme_class_weights() if a.classes == 10 else None
    ol.loss_forward_backward(heads, asg, a.img, ol.LossHParams(), "ciou", 1e-7, w, upstream=(n, n, n))
    det = od.get_detections(heads, levels).numpy()
    res = on.non_max_suppression(det, 0.001, 0.6, backend=nms_backend)
    for r, (gb, gl) in zip(res, targets):
        om.match_image(r.dets, gb.numpy(), gl.numpy(), a.classes)


def _nms_backend():
    try:
        import torchvision  # noqa: F401

        return "torchvision"
    except Exception:
        return "c"


def cpu_baseline(a, n_img):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    backend = _nms_backend()
    heads, targets = _cpu_inputs(a, n_img)
    t0 = time.perf_counter()
    _cpu_step(a, heads, targets, backend)
    dt = time.perf_counter() - t0
    return {"value": n_img / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n_img} images of the step's workload (oracle port: assign + loss fwd/bwd via torch CPU "
                      f"autograd, decode, {backend} greedy NMS conf 0.001 / IoU 0.6, COCO matching), {dt:.1f} s",
            "seconds": dt}


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path.  The reference is pure Python
    (no compiled path, nothing to build into oracle/_ref) and /root/reference does not travel to the GPU
    box, so this times the oracle port: the same ATen ops through torch CPU + torchvision's CPU NMS."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    backend = _nms_backend()
    total = a.steps + a.warmup
    n_img = 2 if total <= 12 else 1  # bounded sample: ~1.5-3 s of CPU work per image (O(n^2) NMS)
    heads, targets = _cpu_inputs(a, n_img)
    for _ in range(min(a.warmup, 1)):  # the CPU path has no warm-up effect worth more than one pass
        _cpu_step(a, heads, targets, backend)
    t0 = time.perf_counter()
    done = 0
    for _ in range(a.steps):
        _cpu_step(a, heads, targets, backend)
        done += 1
        if time.perf_counter() - t0 > 240 and done >= 3:  # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    value = n_img * done / dt
    sample = (f"{n_img} image(s) per step of the {a.batch}-image batch, {done} of {a.steps} requested steps timed "
              f"(oracle port on torch CPU, {backend} NMS)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": done,
        "warmup": min(a.warmup, 1), "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "images_per_step": n_img, "img": a.img, "classes": a.classes},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
