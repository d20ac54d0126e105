#!/usr/bin/env python
"""Benchmark of the YOLOv5s detection tail (BASELINE.json metric: head+loss+NMS images/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path (its files from oracle/_ref)

One "step" = one pass of the hot path over one batch of `--batch` images per GPU (default: config[1] of
BASELINE.json, 64 images/GPU at 640x640, 10 classes, class-weighted loss):
  train tail : target assignment + CIoU/objectness/class-weighted BCE loss, forward AND backward
               (kodb_train_tail, one kernel)
  eval tail  : sigmoid/grid/anchor decode + conf 0.001 filter + class-offset NMS (IoU 0.6, 300/img)
               + per-class COCO matching (kodb_eval_tail, two kernels)
`value` is device-resident throughput: the K timed steps are ONE CUDA-graph launch (device-bound: the host is out
of the picture); per input set an eval chain (candidates -> select/NMS) and an independent train chain, consecutive steps
pipelined over the input sets; CUDA events, max over ranks, with the per-rank spread and the strictly-in-order one-stream
step reported beside it.  `e2e` goes through the public
Python drop-ins with host buffers (H2D of the head maps and targets, D2H of losses/detections/records inside the
timed region).  `extra_configs` holds the other BASELINE configurations (eval B=128, exactly 10 targets/image,
1280x1280 with 80 classes), each with its own roofline block, and the Detect head (one tcgen05 launch,
planar and decode epilogue) against the reference's nine cuDNN convolutions.  `roofline.hbm_probe`: what plain fill / read /
copy kernels reach on this box, alone and concurrently.  Under torchrun rank 0 also recomputes the AP over the
union of all ranks' detections in one process and checks the NCCL-gathered global AP against it (`map_check`).
Prints ONE JSON line on rank 0.
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
# the train tails and the eval tails of the K steps as independent chains of the graph (they share inputs, not results):
# 1.5 us per step better than joining them at every step (tools/exp_sets.sh); KODB_BENCH_SPLIT=0 restores the joined form
SPLIT_CHAINS = os.environ.get("KODB_BENCH_SPLIT", "1") == "1"
# the select/NMS kernels (64 latency-bound CTAs) on streams of their own with the highest stream priority: their CTAs are
# placed ahead of the pending CTAs of the streaming kernels (-0.6 us per step, tools/exp_selprio.sh); 0 = on the candidate
# kernel's stream
SEL_PRIO = int(os.environ.get("KODB_BENCH_SEL_PRIO", "-3"))
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
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations (extra_configs)")
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


def hbm_probe(dev, mb=96, nbuf=4, reps=20):
    """Context for the roofline fractions, measured live with plain torch kernels on buffers of the step's size (rotating
    over `nbuf` buffers > L2): a pure write stream (fill), a pure read stream (sum) and a copy, each alone, and the fill
    and the read as two concurrent streams -- on this GPU a read stream and a write stream share the HBM and do not overlap
    (tools/micro/rw_mix.cu), which is why the step's streaming kernels add up."""
    import torch
    n = mb * (1 << 20) // 4
    bufs = [torch.empty(n, dtype=torch.float32, device=dev).normal_() for _ in range(nbuf)]
    dst = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(nbuf)]
    outs = torch.empty(nbuf, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fa, fb=None):
        best = None
        for _ in range(3):
            torch.cuda.synchronize(dev)
            cur = torch.cuda.current_stream(dev)
            e0.record(cur)
            s1.wait_stream(cur); s2.wait_stream(cur)
            for i in range(reps):
                with torch.cuda.stream(s1):
                    fa(i % nbuf)
                if fb is not None:
                    with torch.cuda.stream(s2):
                        fb(i % nbuf)
            cur.wait_stream(s1); cur.wait_stream(s2)
            e1.record(cur)
            torch.cuda.synchronize(dev)
            t = e0.elapsed_time(e1) / reps
            best = t if best is None else min(best, t)
        return best

    fill = lambda i: dst[i].zero_()
    read = lambda i: torch.sum(bufs[i], dim=0, out=outs[i])
    copy = lambda i: dst[i].copy_(bufs[i])
    b = mb * (1 << 20)
    t_f, t_r, t_c, t_fr = timed(fill), timed(read), timed(copy), timed(fill, read)
    gbs = lambda nbytes, ms: round(nbytes / (ms * 1e-3) / 1e9, 1)
    return {"buffer_mb": mb, "fill_gbs": gbs(b, t_f), "read_gbs": gbs(b, t_r), "copy_gbs_read_plus_write": gbs(2 * b, t_c),
            "fill_and_read_concurrent_gbs_total": gbs(2 * b, t_fr), "fill_us": round(t_f * 1e3, 2), "read_us": round(t_r * 1e3, 2),
            "fill_and_read_concurrent_us": round(t_fr * 1e3, 2),
            "note": "torch kernels (zero_, sum, copy_) on rotating buffers > L2; concurrent = two streams"}


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
def pin_to_gpu_numa_node(local_rank):
    """Bind this rank to the CPUs of its GPU's NUMA node before anything is allocated: pinned host buffers are then
    node-local (first touch) and ranks do not migrate over each other's cores.  Best effort; returns a description."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out  # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": None}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # pragma: no cover
        return {"numa_node": None, "note": str(e)[:80]}


class TailWorkload:
    """One BASELINE configuration resident on the device: NSETS rotating input sets (> L2) with their output buffers and
    workspaces, and the C-ABI calls of a step."""

    def __init__(self, env, B, S, Cn, n_gt=None, nsets=NSETS, train=True, evalt=True, tag=""):
        import torch
        from object_detection_cib_b200 import _host as H
        from object_detection_cib_b200 import nms as N
        from object_detection_cib_b200 import synthetic as syn
        from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape
        self.env, self.B, self.S, self.Cn, self.train, self.evalt = env, B, S, Cn, train, evalt
        dev, lib, rank = env["dev"], env["lib"], env["rank"]
        shape, info = FeatureShape(S, S), env["info"]
        self.loss_mod = env["make_loss"](Cn)
        self.evaluator = env["make_evaluator"](Cn)
        self.sets = []
        for s in range(nsets):
            seed = 2023 + 17 * s + 101 * rank
            heads = syn.make_head(B, S, Cn, seed=seed, device=dev)
            targets = syn.make_targets(B, S, Cn, seed=seed, n_min=0 if n_gt is None else n_gt, n_max=10 if n_gt is None else n_gt,
                                       size_hi=204.0 * (S / 640.0))
            syn.inject_ground_truth(heads, targets, S, fraction=0.3, seed=seed)
            net = [DetectionHeadResult(*lv) for lv in heads]
            plan = self.loss_mod.plan(shape, net, targets)
            # one train workspace per input set: steps of different sets may be in flight at the same time
            plan = plan._replace(workspace=H.workspace(f"train{tag}{s}", dev, plan.workspace.numel()))
            grads = [H.alloc_planar_like(t) for t in plan.heads.tensors] if train else None
            losses = torch.zeros(3, dtype=torch.float32, device=dev)
            nass = torch.zeros(3, dtype=torch.int32, device=dev)
            levels = H.level_cfgs(info, S, S)
            rows = sum(lv.num_anchors * lv.H * lv.W for lv in levels)
            cfg = N._nms_cfg(Cn, 0.001, 0.6, None, N.MAX_DET, N.MAX_NMS)
            res, out = N._alloc_out(B, N.MAX_DET, dev)
            match = N._alloc_match(B, N.MAX_DET, self.evaluator.gt_count, res.counts, dev, N.IOU_THRESHOLDS)
            ws = H.workspace(f"eval{tag}{s}", dev, lib.kodb_eval_workspace_bytes(B, rows, Cn)) if evalt else None
            self.sets.append(dict(heads=heads, targets=targets, net=net, plan=plan, grads=grads, losses=losses, nass=nass,
                                  levels=levels, cfg=cfg, res=res, out=out, match=match, ws=ws, seed=seed,
                                  gstruct=H.grad_structs(grads) if train else None))
        self.up = (C.c_float * 3)(*((float(B),) * 3))
        self.errs = []
        self.launches_per_step = (1 if train else 0) + (2 if evalt else 0)

    def train_call(self, st):
        from object_detection_cib_b200 import _host as H
        p = st["plan"]
        rc = self.env["lib"].kodb_train_tail(p.heads.structs, p.levels, C.byref(p.packed.struct), C.byref(p.cfg),
                                             p.pos_weight.data_ptr() if p.pos_weight is not None else None, self.up, None, 0,
                                             st["gstruct"], st["losses"].data_ptr(), st["nass"].data_ptr(),
                                             p.workspace.data_ptr(), p.workspace.numel(), H.stream_ptr(self.env["dev"]))
        if rc:
            self.errs.append(("train", rc))

    def eval_call(self, st, stage_mask=0):
        from object_detection_cib_b200 import _host as H
        p = st["plan"]
        st["cfg"].stage_mask = stage_mask
        rc = self.env["lib"].kodb_eval_tail(p.heads.structs, st["levels"], self.B, C.byref(st["cfg"]), C.byref(st["out"]),
                                            C.byref(p.packed.struct), C.byref(st["match"].struct), st["ws"].data_ptr(),
                                            st["ws"].numel(), H.stream_ptr(self.env["dev"]))
        st["cfg"].stage_mask = 0
        if rc:
            self.errs.append(("eval", rc))

    def step_serial(self, i):
        st = self.sets[i % len(self.sets)]
        if self.train:
            self.train_call(st)
        if self.evalt:
            self.eval_call(st)

    def step_forked(self, i, side, ev):
        """candidate kernel (all SMs, HBM-bound) first; then the select/NMS kernel (one CTA per image, latency bound) on
        the current stream and the train tail (HBM-bound stream) on `side`: they share the GPU."""
        import torch
        st = self.sets[i % len(self.sets)]
        if not (self.train and self.evalt):
            return self.step_serial(i)
        dev = self.env["dev"]
        cur = torch.cuda.current_stream(dev)
        self.eval_call(st, 1)
        ev.record(cur)
        self.eval_call(st, 2)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            self.train_call(st)
        cur.wait_stream(side)

    def algorithmic_bytes(self):
        ac = sum(3 * (self.S // s) ** 2 for s in (8, 16, 32))
        b_train = 4 * ac * (1 + 5 + self.Cn) if self.train else 0
        b_eval = 4 * ac * (5 + self.Cn) + 300 * (6 * 4 + 8) if self.evalt else 0
        return ac, b_train, b_eval


def capture_steps(wl, n_steps, mode, dev):
    """ONE CUDA graph holding `n_steps` consecutive steps.
      "pipelined": the steps of input set s run on branch s, so up to NSETS independent steps are in flight and the
                   latency-bound select/NMS kernel of one step overlaps the streaming kernels of the next; by default
                   (SPLIT_CHAINS) the eval tails (candidates -> select/NMS) and the train tails of a set form two independent
                   chains, otherwise the train tail forks off behind the candidate kernel and joins before the set's next step;
      "step":      steps strictly one after the other; inside a step the train tail runs beside the select/NMS kernel
                   (both only read the head maps) -- what a loop that cannot overlap steps sees;
      "serial":    every kernel strictly after the previous one.
    A single graph launch makes the timed region device-bound: the host is out of the picture after one call."""
    import torch
    ns = len(wl.sets)
    cap = torch.cuda.Stream(device=dev)
    branches = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(ns)]
    sides = [torch.cuda.Stream(device=dev) for _ in range(ns)]
    evs = [torch.cuda.Event() for _ in range(ns)]
    sels = [torch.cuda.Stream(device=dev, priority=SEL_PRIO) for _ in range(ns)] if SEL_PRIO else []
    sel_evs = [torch.cuda.Event() for _ in range(ns)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        if mode == "pipelined" and SPLIT_CHAINS and wl.train and wl.evalt:
            # the train tails and the eval tails of the K steps as independent chains (they share inputs, not results)
            cur = torch.cuda.current_stream(dev)
            for b in branches + sides + sels:
                b.wait_stream(cur)
            for i in range(n_steps):
                st = wl.sets[i % ns]
                if SEL_PRIO:
                    # the select/NMS kernel on a stream of its own with a higher priority than the candidate kernel
                    with torch.cuda.stream(branches[i % ns]):
                        if i >= ns:
                            branches[i % ns].wait_event(sel_evs[i % ns])  # the set's previous select kernel (workspace reuse)
                        wl.eval_call(st, 1)
                        evs[i % ns].record(branches[i % ns])
                    with torch.cuda.stream(sels[i % ns]):
                        sels[i % ns].wait_event(evs[i % ns])
                        wl.eval_call(st, 2)
                        sel_evs[i % ns].record(sels[i % ns])
                else:
                    with torch.cuda.stream(branches[i % ns]):
                        wl.eval_call(st)
                with torch.cuda.stream(sides[i % ns]):
                    wl.train_call(st)
            for b in branches + sides + sels:
                cur.wait_stream(b)
        elif mode == "pipelined":
            cur = torch.cuda.current_stream(dev)
            for b in branches:
                b.wait_stream(cur)
            for i in range(n_steps):
                with torch.cuda.stream(branches[i % ns]):
                    wl.step_forked(i, sides[i % ns], evs[i % ns])
            for b in branches:
                cur.wait_stream(b)
        elif mode == "step":
            for i in range(n_steps):
                wl.step_forked(i, sides[0], evs[0])
        else:
            for i in range(n_steps):
                wl.step_serial(i)
    return g, (cap, branches, sides, evs, sels, sel_evs)


def measure_workload(wl, steps, warmup, env, full=True):
    """value (K pipelined steps, one graph launch), the one-stream step, per-kernel device times."""
    import torch
    import torch.distributed as dist
    dev, world = env["dev"], env["world"]
    ns = len(wl.sets)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(ns):  # first launches set per-device function attributes (not capturable)
        wl.step_serial(i)
    torch.cuda.synchronize(dev)
    g_pipe, keep1 = capture_steps(wl, steps, "pipelined", dev)
    g_warm, keep2 = capture_steps(wl, max(warmup, 3), "pipelined", dev)
    g_warm.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_load0 = time.time()
    e0.record()
    g_pipe.replay()  # EXACTLY `steps` steps
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t_all = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = [ms]
    if world > 1:
        gathered = [torch.zeros_like(t_all) for _ in range(world)]
        dist.all_gather(gathered, t_all)
        per_rank = [float(t.item()) for t in gathered]
    ms_total = max(per_rank)
    out = {"ms_total": ms_total, "ms_per_step": ms_total / steps, "value": world * wl.B * steps / (ms_total / 1e3),
           "per_rank_ms_per_step": {"min": min(per_rank) / steps, "median": statistics.median(per_rank) / steps,
                                    "max": max(per_rank) / steps}, "t_load0": t_load0}
    # ---- the same K steps without overlap between steps (what a train loop that cannot overlap steps sees), and with
    # every kernel strictly after the previous one
    for mode, key in (("step", "one_stream_ms_per_step"), ("serial", "serial_kernels_ms_per_step")):
        g_one, keep3 = capture_steps(wl, steps, mode, dev)
        g_one.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        g_one.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        out[key] = e0.elapsed_time(e1) / steps
        del g_one, keep3
    if not full:
        out["t_load1"] = time.time()
        return out
    # ---- per-kernel device time: back-to-back eager launches of the same work between ONE pair of events, queued behind a
    # spin kernel (the host needs longer to issue a launch through ctypes than these kernels run); the candidate / select
    # split of the eval pair from an event-separated pass
    n_inst = min(max(steps, 20), 60)
    t_train = t_pair = t_cand_ev = t_sel_ev = t_train_ev = None
    if wl.train and wl.evalt:
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_inst)]
        torch.cuda._sleep(int(0.04 * 1.9e9))
        for i in range(n_inst):
            st = wl.sets[i % ns]
            evs[i][0].record(); wl.train_call(st); evs[i][1].record(); wl.eval_call(st, 1); evs[i][2].record()
            wl.eval_call(st, 2); evs[i][3].record()
        torch.cuda.synchronize(dev)
        t_train_ev = statistics.median(e[0].elapsed_time(e[1]) for e in evs)
        t_cand_ev = statistics.median(e[1].elapsed_time(e[2]) for e in evs)
        t_sel_ev = statistics.median(e[2].elapsed_time(e[3]) for e in evs)
    bb = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda._sleep(int(0.04 * 1.9e9))
    bb[0].record()
    if wl.train:
        for i in range(n_inst):
            wl.train_call(wl.sets[i % ns])
    bb[1].record()
    if wl.evalt:
        for i in range(n_inst):
            wl.eval_call(wl.sets[i % ns])
    bb[2].record()
    torch.cuda.synchronize(dev)
    if wl.train:
        t_train = bb[0].elapsed_time(bb[1]) / n_inst
    if wl.evalt:
        t_pair = bb[1].elapsed_time(bb[2]) / n_inst
        if t_cand_ev is None:  # eval-only workload: event-separated split
            evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_inst)]
            torch.cuda._sleep(int(0.04 * 1.9e9))
            for i in range(n_inst):
                st = wl.sets[i % ns]
                evs[i][0].record(); wl.eval_call(st, 1); evs[i][1].record(); wl.eval_call(st, 2); evs[i][2].record()
            torch.cuda.synchronize(dev)
            t_cand_ev = statistics.median(e[0].elapsed_time(e[1]) for e in evs)
            t_sel_ev = statistics.median(e[1].elapsed_time(e[2]) for e in evs)
    out["t_load1"] = time.time()
    out.update(t_train=t_train, t_pair=t_pair, t_cand_ev=t_cand_ev, t_sel_ev=t_sel_ev, t_train_ev=t_train_ev)
    if wl.errs:
        raise RuntimeError(f"C-ABI calls failed: {wl.errs[:3]}")
    return out


def roofline_blocks(wl, m, peak):
    """Per-kernel and per-step achieved GB/s against the measured HBM peak (SURVEY 8d algorithmic bytes)."""
    ac, b_train, b_eval = wl.algorithmic_bytes()
    B, Cn = wl.B, wl.Cn
    b_cand = 4 * ac * (1 + Cn) if wl.evalt else 0
    b_sel = b_eval - b_cand
    stages, stages_ms = {}, {}
    gbs = lambda nbytes, ms: nbytes * B / (ms * 1e-3) / 1e9
    if wl.train:
        stages_ms["train_tail_kernel"] = m["t_train"]
        stages["train_tail"] = {"achieved_gbs": gbs(b_train, m["t_train"]), "frac": gbs(b_train, m["t_train"]) / peak}
    if wl.evalt:
        t_cand = m["t_pair"] * m["t_cand_ev"] / (m["t_cand_ev"] + m["t_sel_ev"])
        t_sel = m["t_pair"] - t_cand
        stages_ms["cand_floor_kernel"], stages_ms["select_nms_kernel"] = t_cand, t_sel
        stages["cand_floor_kernel"] = {"achieved_gbs": gbs(b_cand, t_cand), "frac": gbs(b_cand, t_cand) / peak}
        stages["select_nms_kernel"] = {"bound": "latency (one CTA per image)", "achieved_gbs": gbs(b_sel, t_sel),
                                       "frac": gbs(b_sel, t_sel) / peak}
        stages["eval_tail"] = {"achieved_gbs": gbs(b_eval, m["t_pair"]), "frac": gbs(b_eval, m["t_pair"]) / peak}
    stages["step"] = {"achieved_gbs": gbs(b_train + b_eval, m["ms_per_step"]), "frac": gbs(b_train + b_eval, m["ms_per_step"]) / peak}
    stages["step_one_stream"] = {"achieved_gbs": gbs(b_train + b_eval, m["one_stream_ms_per_step"]),
                                 "frac": gbs(b_train + b_eval, m["one_stream_ms_per_step"]) / peak}
    # the roofline object describes the dominant HBM-bound kernel: of the streaming kernels the one with the longer launch
    cands = [k for k in ("train_tail_kernel", "cand_floor_kernel") if k in stages_ms]
    dom = max(cands, key=lambda k: stages_ms[k])
    bytes_dom = (b_train if dom == "train_tail_kernel" else b_cand) * B
    ach = bytes_dom / (stages_ms[dom] * 1e-3) / 1e9
    return dom, bytes_dom, ach, stages, stages_ms


EXTRA_CONFIGS = {
    # BASELINE.json configs[2], [3], [4]; configs[1] is the headline workload
    "cfg3_eval128": dict(B=128, S=640, Cn=10, n_gt=None, train=False, evalt=True,
                         name="eval tail only, 128 img/GPU 640x640, 10 classes: decode + conf 0.001 dense-candidate NMS + mAP matching"),
    "cfg4_dense10": dict(B=64, S=640, Cn=10, n_gt=10, train=True, evalt=False,
                         name="train tail only, 64 img/GPU 640x640, exactly 10 targets/image (mosaic/mixup-like): build_targets stress"),
    "cfg5_coco1280": dict(B=8, S=1280, Cn=80, n_gt=None, train=True, evalt=True,
                          name="train + eval tail, 8 img/GPU 1280x1280, 80 classes (large grids, top-30000 cut)"),
}


def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    numa = pin_to_gpu_numa_node(local)

    from object_detection_cib_b200 import _lib as L
    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200 import synthetic as syn
    from object_detection_cib_b200.iou import IoUCalculator, IoUType
    from object_detection_cib_b200.label_assignment import Yolov5LabelAssigner
    from object_detection_cib_b200.loss import Yolov5Loss, Yolov5LossParams
    from object_detection_cib_b200.map_eval import DeviceCocoMapEvaluator
    from object_detection_cib_b200.types import DetectionHeadResult, FeatureShape, default_anchor_info

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    info = default_anchor_info()
    env = dict(dev=dev, lib=L.lib(), rank=rank, world=world, info=info,
               make_loss=lambda Cn: Yolov5Loss(Yolov5LabelAssigner(info), Yolov5LossParams.get_default(),
                                               IoUCalculator(IoUType.ciou, 1e-7), syn.readme_class_weights() if Cn == 10 else None),
               make_evaluator=lambda Cn: DeviceCocoMapEvaluator(Cn, device=dev))
    B, S, Cn = a.batch, a.img, a.classes
    shape = FeatureShape(S, S)
    wl = TailWorkload(env, B, S, Cn)
    loss_mod, sets = wl.loss_mod, wl.sets

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    m = measure_workload(wl, a.steps, max(a.warmup, 3), env)
    clocks = sampler.stop(m["t_load0"], m["t_load1"]) if rank == 0 else None

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

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        n_e2e = max(3, min(a.steps, 30))
        e2e_run(3)
        barrier()
        t0 = time.perf_counter()
        e2e_run(n_e2e)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        per_rank = [dt]
        if world > 1:
            gathered = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(gathered, tt)
            per_rank = [float(t.item()) for t in gathered]
        dmax = max(per_rank)
        e2e = {"value": world * B * n_e2e / dmax, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "steps": n_e2e,
               "h2d_gbs_per_rank": {"min": h2d * n_e2e / max(per_rank) / 1e9, "max": h2d * n_e2e / min(per_rank) / 1e9},
               "api": "Yolov5Loss.forward+backward, nms.eval_tail (pinned host head maps -> device -> pinned host results; H2D of step i+1 overlaps step i)"}
        del dev_in, host_heads

    # ---- epoch-end exchange (the only collective of the path): gather records + global AP ---------------
    evaluator = wl.evaluator
    evaluator.reset()  # the timed loops accumulated GT counts for thousands of steps; start a clean "epoch"
    for st in sets:
        st["match"] = N._alloc_match(B, N.MAX_DET, evaluator.gt_count, st["res"].counts, dev, N.IOU_THRESHOLDS)
        wl.eval_call(st)
        evaluator.add_records(st["match"].records)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    ap_global = evaluator.average_precision()  # NCCL all_gather of the records + all_reduce of the GT counts when world > 1
    report = evaluator.get_report()
    epoch_end_ms = (time.perf_counter() - t0) * 1e3
    # Correctness of the multi-GPU result: rank 0 recomputes the AP in ONE process over the union of every rank's raw
    # detections and ground truth (detections gathered, targets regenerated from the ranks' seeds) and compares the tables.
    map_check = None
    if world > 1:
        dets_all = [torch.empty((len(sets) * B, N.MAX_DET, 6), dtype=torch.float32, device=dev) for _ in range(world)]
        cnts_all = [torch.empty((len(sets) * B,), dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(dets_all, torch.cat([st["res"].dets for st in sets], 0).contiguous())
        dist.all_gather(cnts_all, torch.cat([st["res"].counts for st in sets], 0).contiguous())
        if rank == 0:
            union = DeviceCocoMapEvaluator(Cn, device=dev)
            for r in range(world):
                for s in range(len(sets)):
                    seed = 2023 + 17 * s + 101 * r
                    tg = syn.make_targets(B, S, Cn, seed=seed, n_min=0, n_max=10, size_hi=204.0 * (S / 640.0))
                    union.add_detections(dets_all[r][s * B:(s + 1) * B], cnts_all[r][s * B:(s + 1) * B], tg)
            ap_union = union.average_precision(gather=False)
            same_nan = bool((torch.isnan(ap_union) == torch.isnan(ap_global)).all())
            diff = float((torch.nan_to_num(ap_union) - torch.nan_to_num(ap_global)).abs().max())
            map_check = {"global_vs_single_process_union_max_abs_diff": diff, "ok": bool(same_nan and diff < 1e-9),
                         "images": world * len(sets) * B}
            if not map_check["ok"]:
                raise RuntimeError(f"gathered global AP differs from the single-process AP over the union: {map_check}")

    # diagnostic: share of the candidate kernel's warp tiles that listed ALL their candidates in the last eval call
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

    peak, peak_src = measured_peak_gbs()
    dom, bytes_dom, ach, stages, stages_ms = roofline_blocks(wl, m, peak)

    # ---- the other BASELINE configurations (rank 0 of a single-GPU run: the driver's BENCH line) ---------------------
    extra = {}
    if world == 1 and not a.no_extra:
        del wl, sets
        torch.cuda.empty_cache()
        for key, c in EXTRA_CONFIGS.items():
            try:
                w2 = TailWorkload(env, c["B"], c["S"], c["Cn"], n_gt=c["n_gt"], train=c["train"], evalt=c["evalt"], tag=key)
                m2 = measure_workload(w2, min(a.steps, 40), 3, env)
                d2, bd2, ach2, st2, sms2 = roofline_blocks(w2, m2, peak)
                extra[key] = {"workload": c["name"], "images_per_gpu": c["B"], "img": c["S"], "classes": c["Cn"],
                              "steps": min(a.steps, 40), "ms_per_step": m2["ms_per_step"], "value": m2["value"], "unit": UNIT,
                              "one_stream_ms_per_step": m2["one_stream_ms_per_step"],
                              "roofline": {"bound": "hbm", "kernel": d2, "achieved": ach2, "peak": peak, "unit": "GB/s",
                                           "frac": ach2 / peak, "algorithmic_bytes_per_launch": bd2, "kernel_ms": sms2[d2]},
                              "stages_ms": sms2, "roofline_stages": st2}
                del w2
                torch.cuda.empty_cache()
            except Exception as e:  # pragma: no cover
                extra[key] = {"workload": c["name"], "error": str(e)[:200]}
        for key, (hb, hs, hc) in {"f1_detect_head": (B, S, Cn), "f1_detect_head_1280_c80": (8, 1280, 80)}.items():
            try:
                extra[key] = measure_head(hb, hs, hc, peak)
                torch.cuda.empty_cache()
            except Exception as e:  # pragma: no cover
                extra[key] = {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        ac = sum(3 * (S // s) ** 2 for s in (8, 16, 32))
        b_eval = 4 * ac * (5 + Cn) + 300 * 32
        traffic = None
        try:  # dram__bytes_read + dram__bytes_write per launch from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("images_per_launch") == B:
                traffic = tr["kernels"].get(dom)
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "images_per_gpu": B, "global_batch": B * world, "img": S,
                       "classes": Cn, "parallelism": f"dp{world} (shard by image, no data-path collective)",
                       "l2": f"inputs rotate over {NSETS} sets ({NSETS}x{(b_eval * B) / 1e6:.0f} MB head + grads) > 126 MB L2",
                       "launch": f"ONE cuda-graph launch holds all {a.steps} timed steps (device-bound timed region); per step: "
                                 f"candidate kernel -> select/NMS kernel (stream priorities -1 / {SEL_PRIO if SEL_PRIO else -1}) on an eval chain, "
                                 f"train tail (priority 0) on a train chain; "
                                 + ("the eval chain and the train chain of an input set run independently (they share inputs, not "
                                    "results); " if SPLIT_CHAINS else "the chains of a step join before the next step of the set; ")
                                 + f"consecutive steps pipelined over {NSETS} input sets (one pair of chains per set)",
                       "numa": numa},
            "gpu_launches": wl_launches(a.steps),
            "per_rank_ms_per_step": m["per_rank_ms_per_step"],
            "one_stream_ms_per_step": m["one_stream_ms_per_step"],
            "serial_kernels_ms_per_step": m["serial_kernels_ms_per_step"],
            "one_stream_note": "one_stream: steps strictly one after the other (no step pipelining), train tail beside the "
                               "select/NMS kernel inside a step; serial_kernels: every kernel strictly after the previous one",
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                         "traffic_note": "dram bytes under ncu are below the algorithmic bytes for the train tail: the 126 MB L2 "
                                         "defers the write-back of its streaming stores past the end of the kernel",
                         "algorithmic_bytes_per_launch": bytes_dom, "kernel_ms": stages_ms[dom]},
            "stages_ms": dict(stages_ms, note="device time per launch from back-to-back eager launches between one pair of CUDA "
                                              "events, queued behind a spin kernel (train tail alone; eval tail as candidate+select "
                                              "pairs, split in the ratio of the event-separated times)",
                              event_separated_ms={"train_tail_kernel": m["t_train_ev"], "cand_floor_kernel": m["t_cand_ev"],
                                                  "select_nms_kernel": m["t_sel_ev"]}),
            "roofline_stages": stages,
            "e2e": e2e, "clocks": clocks, "epoch_end_ms": epoch_end_ms, "tiles_listed_in_full": full_tiles,
            "map_report": {k: report[k] for k in ("avg_mAP", "mAP_50")}, "map_check": map_check,
            "extra_configs": extra,
        }
        if world == 1 and not a.no_extra:
            try:
                line["roofline"]["hbm_probe"] = hbm_probe(dev)
            except Exception as e:  # pragma: no cover
                line["roofline"]["hbm_probe"] = {"error": str(e)[:200]}
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, a.cpu_sample)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_head(B, S, Cn, peak, n=20, nsets=3):
    """the Detect head (three 1x1 convolutions per level + Rearrange) as ONE tcgen05 launch
    (`kodb_head_forward`) against the reference's own execution -- nine cuDNN convolutions through torch -- on the same
    inputs; device time per call, n calls over `nsets` rotating input sets (> L2 together) in one CUDA graph."""
    import torch
    import torch.nn.functional as F

    from object_detection_cib_b200 import nms as N
    from object_detection_cib_b200.heads import Yolov5Head, fused_heads_detections, fused_heads_forward
    from object_detection_cib_b200.types import FeatureShape, default_anchor_info

    dev = torch.device("cuda", torch.cuda.current_device())
    cfg = ((128, 8), (256, 16), (512, 32))
    torch.manual_seed(0)
    heads = [Yolov5Head(K, 3, Cn, s).to(dev) for K, s in cfg]
    xsets = [[torch.randn(B, K, S // s, S // s, device=dev) for K, s in cfg] for _ in range(nsets)]

    def ref(xs):
        return [[F.conv2d(x, m.conv.weight, m.conv.bias) for m in (h.box_head, h.obj_head, h.cls_head)] for h, x in zip(heads, xs)]

    def timed(fn):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for i in range(3):
                fn(xsets[i % nsets])
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for i in range(n):
                    fn(xsets[i % nsets])
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    old = torch.backends.cudnn.allow_tf32
    try:
        with torch.no_grad():
            t_ours = timed(lambda xs: fused_heads_forward(heads, xs))
            torch.backends.cudnn.allow_tf32 = True
            t_tf32 = timed(ref)
            torch.backends.cudnn.allow_tf32 = False
            t_fp32 = timed(ref)
            info, shape = default_anchor_info(), FeatureShape(S, S)
            t_det = timed(lambda xs: fused_heads_detections(heads, xs, shape, info))
            t_two = timed(lambda xs: N.get_detections(shape, fused_heads_forward(heads, xs), info))
            got, want = fused_heads_forward(heads, xsets[0]), ref(xsets[0])
            err = max(float((g.permute(0, 1, 4, 2, 3).reshape(w.shape) - w).abs().max()) for gl, wl in zip(got, want) for g, w in zip(gl, wl))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    n_out = 3 * (5 + Cn)
    cells = sum((S // s) ** 2 for _, s in cfg)
    b_in = sum(4 * B * K * (S // s) ** 2 for K, s in cfg)
    b_out = 4 * B * n_out * cells
    ach = (b_in + b_out) / (t_ours * 1e-3) / 1e9
    return {"workload": f"Detect head, {B} img {S}x{S}, {Cn} classes: box/obj/cls 1x1 convolutions of the three pyramid levels "
                        f"(K = 128/256/512 -> {n_out} channels) + planar Rearrange, fp32 in, tf32 multiply, fp32 accumulate",
            "kernel": "head_forward_kernel (tcgen05.mma kind::tf32, TMEM accumulators, TMA loads; one launch for all levels)",
            "ms_per_call": t_ours, "value": B / (t_ours * 1e-3), "unit": UNIT, "gpu_launches_per_call": 1,
            "cudnn_tf32_ms_per_call": t_tf32, "cudnn_fp32_ms_per_call": t_fp32, "cudnn_launches_per_call": 9,
            "speedup_vs_cudnn_tf32": t_tf32 / t_ours, "speedup_vs_cudnn_fp32": t_fp32 / t_ours,
            "max_abs_err_vs_fp32_conv": err,
            "eval_side": {"what": "features -> get_detections' (B, AC, 5+C) rows in ONE launch (kodb_head_detections: the same GEMM "
                                  "with the decode epilogue; the planar logits are never written)",
                          "ms_per_call": t_det, "achieved_gbs": (b_in + b_out) / (t_det * 1e-3) / 1e9,
                          "frac": (b_in + b_out) / (t_det * 1e-3) / 1e9 / peak,
                          "planar_head_then_decode_kernel_ms_per_call": t_two, "speedup_vs_two_launches": t_two / t_det,
                          "speedup_vs_cudnn_tf32_convs_alone": t_tf32 / t_det},
            "roofline": {"bound": "hbm", "kernel": "head_forward_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "algorithmic_bytes_per_launch": b_in + b_out,
                         "note": "4*K bytes in + 4*3*(5+C) bytes out per cell; ~17 flop/B at K=128: HBM-bound, the tensor "
                                 "cores only have to keep up"}}


def wl_launches(steps):
    return 3 * steps  # train_tail_kernel, cand_floor_kernel, select_nms_kernel per step


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
    """The reference's CPU path on a bounded sample of the step's workload: its own files from oracle/_ref when staged
    (kind "reference"), else the oracle port (kind "port")."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    backend = _nms_backend()
    heads, targets = _cpu_inputs(a, n_img)
    R = None
    try:
        R = _ref_modules()
    except Exception:  # pragma: no cover
        R = None
    t0 = time.perf_counter()
    if R:
        _ref_step(a, R, heads, targets)
    else:
        _cpu_step(a, heads, targets, backend)
    dt = time.perf_counter() - t0
    what = ("the reference's own Yolov5Loss fwd+bwd, get_detections, non_max_suppression (oracle/_ref) on torch CPU; COCO "
            "matching by the oracle's restatement") if R else \
           f"oracle port: assign + loss fwd/bwd via torch CPU autograd, decode, {backend} greedy NMS conf 0.001 / IoU 0.6, COCO matching"
    return {"value": n_img / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference" if R else "port",
            "sample": f"{n_img} images of the step's workload ({what}), {dt:.1f} s", "seconds": dt}


def _ref_modules():
    """The reference's own modules from oracle/_ref (staged by build(), see oracle/ref_stage.py), or None."""
    from oracle import ref_stage

    if not ref_stage.staged() or not ref_stage.install(ref_stage.STAGED_ZIP):
        return None
    from kod.core.anchors.info import AnchorBoxInfo
    from kod.core.bbox.iou import IoUCalculator, IoUType
    from kod.core.label_assignment.yv5 import AssignmentAnchorInfo, Yolov5LabelAssigner
    from kod.core.nms import non_max_suppression
    from kod.core.types import FeatureShape
    from kod.data.detection import DetectionTarget
    from kod.lightning.experiments.yv5_baseline.exp import DefaultYolov5Experiment
    from kod.lightning.experiments.yv5_baseline.loss import Yolov5Loss, Yolov5LossParams
    from kod.lightning.experiments.yv5_baseline.type_defs import LayerwiseAnchorInfo
    from kod.nn.heads.types import DetectionHeadResult
    from kod.nn.networks.yolov5 import Yolov5NetworkResult

    return dict(locals())


def _ref_step(a, R, heads, targets):
    """One pass of the REAL reference over `heads`: Yolov5Loss forward + backward (loss.py:46-248, exp.py:124-130),
    get_detections (exp.py:70-102), non_max_suppression (nms.py:9-75); the COCO matching is third-party code that is
    absent here (pycocotools behind vision_evaluation): the oracle's restatement stands in for it."""
    import torch

    from object_detection_cib_b200 import synthetic as syn
    from oracle import coco_map as om

    S, n = a.img, heads[0][0].shape[0]
    infos = [R["AnchorBoxInfo"](stride=s, boxes_wh=[R["FeatureShape"](width=w, height=h) for (w, h) in an])
             for s, an in zip(syn.STRIDES, syn.ANCHORS_PX)]
    w = syn.readme_class_weights() if a.classes == 10 else None
    loss_mod = R["Yolov5Loss"](R["Yolov5LabelAssigner"](R["AssignmentAnchorInfo"](*infos)), R["Yolov5LossParams"].get_default(),
                               R["IoUCalculator"](R["IoUType"]("ciou"), eps=1e-7), weights=w)
    if loss_mod.weights is not None:
        loss_mod.weights = loss_mod.weights.cpu()  # loss.py:58-61 moves them to cuda when one is visible: this is the CPU arm
    leaves = R["Yolov5NetworkResult"](*[R["DetectionHeadResult"](*[t.detach().clone().requires_grad_(True) for t in lv])
                                        for lv in heads])
    tg = tuple(R["DetectionTarget"](boxes=b, labels=l) for b, l in targets)
    shape = R["FeatureShape"](width=S, height=S)
    res = loss_mod(shape, leaves, tg)
    (n * (res.localization + res.classification + res.objectness)).backward()

    class _Exp:  # only what get_detections touches
        anchor_info = R["LayerwiseAnchorInfo"](*infos)
        get_detections = R["DefaultYolov5Experiment"].get_detections

    with torch.no_grad():
        det = _Exp().get_detections(shape, R["Yolov5NetworkResult"](*[R["DetectionHeadResult"](*lv) for lv in heads]))
        out = R["non_max_suppression"](det, conf_thres=0.001, nms_thres=0.6)
    for r, (gb, gl) in zip(out, targets):
        om.match_image(r.numpy().reshape(-1, 6), gb.numpy(), gl.numpy(), a.classes)


def run_reference(a):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores -- its unmodified
    python files from oracle/_ref (staged by build() where /root/reference exists; the reference has no compiled code),
    else the oracle port (same ATen ops through torch CPU + torchvision's CPU NMS)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    backend = _nms_backend()
    R = None
    try:
        R = _ref_modules()
    except Exception as e:  # pragma: no cover - a broken staging falls back to the port, and says so
        print(f"[bench] oracle/_ref not usable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    kind = "reference" if R else "port"
    step = (lambda h, t: _ref_step(a, R, h, t)) if R else (lambda h, t: _cpu_step(a, h, t, backend))
    total = a.steps + a.warmup
    n_img = 2 if total <= 12 else 1  # bounded sample: ~1.5-3 s of CPU work per image (O(n^2) NMS)
    heads, targets = _cpu_inputs(a, n_img)
    for _ in range(min(a.warmup, 1)):  # the CPU path has no warm-up effect worth more than one pass
        step(heads, targets)
    t0 = time.perf_counter()
    done = 0
    for _ in range(a.steps):
        step(heads, targets)
        done += 1
        if time.perf_counter() - t0 > 240 and done >= 3:  # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    value = n_img * done / dt
    what = ("the reference's own Yolov5Loss fwd+bwd, get_detections and non_max_suppression from oracle/_ref on torch CPU; "
            "COCO matching by the oracle's restatement of pycocotools (third party, absent)") if R else \
           f"oracle port on torch CPU, {backend} NMS"
    sample = f"{n_img} image(s) per step of the {a.batch}-image batch, {done} of {a.steps} requested steps timed ({what})"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": done,
        "warmup": min(a.warmup, 1), "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "images_per_step": n_img, "img": a.img, "classes": a.classes},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
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
