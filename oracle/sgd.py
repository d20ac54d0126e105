"""Oracle (test infrastructure): numpy restatement of the optimizer step + warm-up of the reference's train loop.

* `sgd_step` – torch.optim.SGD's `_single_tensor_sgd` (third party: torch, installed 2.11.0; reference pins 2.0.1) as
  configured by kod/configs/nn/optimizers/smart_sgd.yaml (momentum 0.937, nesterov, lr 0.01) and grouped by
  SmartOptimizer (kod/nn/optim/smart.py:11-60): d = g + wd*p; buf = mu*buf + d (buf = d on the first step);
  d = d + mu*buf (nesterov); p = p - lr*d.  fp32 with the fused multiply-adds of ATen's add(alpha=) kernels.
* `warmup_hyper` – OptimizerWarmupUpdater.__call__ (kod/lightning/experiments/yv5_baseline/warmup.py:21-58) with the
  scheduler factor sch_linear (kod/nn/optim/schedulers.py:18-19).
Pinned by tests/golden/sgd_warmup.npz (the real reference classes run on a small net for 7 steps).
"""
from __future__ import annotations

import math

import numpy as np


def _fma(a, b, c):
    """fp32 fused multiply-add: exact product + sum in fp64 (24+24 bits fit in 53), one rounding to fp32."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def sgd_step(p, g, buf, lr, momentum, weight_decay, nesterov=True):
    """One step for one tensor; returns (p, buf).  `buf` None = first step."""
    p, g = p.astype(np.float32), g.astype(np.float32)
    f = lambda v: np.full_like(p, np.float32(v))
    d = g
    if weight_decay != 0:
        d = _fma(f(weight_decay), p, d)
    if momentum != 0:
        buf = d.copy() if buf is None else ((f(momentum) * buf).astype(np.float32) + d).astype(np.float32)
        d = _fma(f(momentum), buf, d) if nesterov else buf
    return _fma(f(-lr), d, p), buf


def sch_linear(x, max_epochs, lrf):
    return (1 - x / max_epochs) * (1.0 - lrf) + lrf


def warmup_hyper(step, epoch, max_warmup_steps, group_name, initial_lr, warmup_bias_lr, warmup_momentum, momentum, sch_fn):
    xi = [0, max_warmup_steps]
    lr = np.interp(step, xi, [warmup_bias_lr if group_name == "bias_params" else 0.0, initial_lr * sch_fn(epoch)])
    mu = np.interp(step, xi, [warmup_momentum, momentum])
    return float(lr), float(mu)
