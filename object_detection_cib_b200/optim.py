"""Optimizer step + warm-up of the reference's train loop on the CUDA kernel `kodb_sgd_step`
(csrc/sgd.cu).

* `ArenaSGD` – `torch.optim.SGD(lr, momentum, dampening, weight_decay, nesterov)` semantics
  (`_single_tensor_sgd`, op for op) over FLAT arenas: at the first step every parameter, its gradient and its
  momentum buffer become views of three contiguous fp32 buffers laid out group by group, and a step is ONE streaming
  kernel launch (plus the reset of the gradients for the next backward) instead of a few hundred small per-tensor
  kernels.  `param_groups`, `state[p]["momentum_buffer"]`, `state_dict()` keep torch's shapes, so LR schedulers, the
  warm-up updater and checkpoints work unchanged.
* `SmartOptimizer` – kod/nn/optim/smart.py:11-60: the three named groups (`bias_params` without decay, `decay_params`,
  `norm_params` without decay) around any optimizer factory.
* `OptimizerWarmupUpdater` – kod/lightning/experiments/yv5_baseline/warmup.py:11-58: per-step linear interpolation of
  the groups' lr (bias: warmup_bias_lr -> lr, others: 0 -> lr) and momentum during warm-up.

No CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _host as H
from . import _lib as L


def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


class ArenaSGD(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0, weight_decay: float = 0.0,
                 nesterov: bool = False, zero_grad_in_step: bool = False):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError("lr, momentum and weight_decay must be non-negative")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")  # torch.optim.SGD's own check
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))
        #: True: the step kernel also writes grads = 0 (the next backward accumulates into zeros) and `zero_grad()`
        #: becomes a no-op -- saves one pass over the gradients per step.  Contract: gradients are then only ever
        #: discarded by stepping.  False (default): `zero_grad()` clears the arena with one memset, like torch.
        self.zero_grad_in_step = bool(zero_grad_in_step)
        self._arena = None

    # ---- arena -----------------------------------------------------------------------------------------
    def _signature(self):
        return tuple(tuple(id(p) for p in g["params"]) for g in self.param_groups)

    def _build(self) -> None:
        groups = self.param_groups
        if len(groups) > 8:
            raise ValueError("at most 8 parameter groups")
        ps = [p for g in groups for p in g["params"]]
        if not ps:
            raise ValueError("optimizer got no parameters")
        dev = ps[0].device
        for p in ps:
            H.require_cuda(p, "parameter")
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("ArenaSGD needs float32 parameters on one CUDA device (the reference trains in fp32)")
        nest = {bool(g["nesterov"]) for g in groups}
        if len(nest) != 1:
            raise ValueError("nesterov must be the same for every group")
        seg, offs, total = [0], {}, 0
        for g in groups:
            for p in g["params"]:
                offs[id(p)] = total
                total += _round4(p.numel())  # every tensor starts 16-byte aligned
            seg.append(total)
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p in ps:
                o, n = offs[id(p)], p.numel()
                view = flat_p[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                gview = flat_g[o:o + n].view(p.shape)
                if p.grad is not None:
                    gview.copy_(p.grad)
                p.grad = gview
                st = self.state[p]
                mview = flat_m[o:o + n].view(p.shape)
                if st.get("momentum_buffer") is not None:
                    mview.copy_(st["momentum_buffer"])
                st["momentum_buffer"] = mview
        self._arena = dict(p=flat_p, g=flat_g, m=flat_m, offs=offs, seg=(C.c_int64 * len(seg))(*seg), sig=self._signature(),
                           hyper=torch.zeros((len(groups), 4), dtype=torch.float32, device=dev), dev=dev,
                           nesterov=nest.pop(), staging=[torch.zeros((len(groups), 4), dtype=torch.float32).pin_memory()
                                                         for _ in range(8)], events=[None] * 8, tick=0, last=None)

    def _adopt_stray_gradients(self) -> None:
        """A gradient that is not the arena view (someone called zero_grad(set_to_none=True) on the module, or assigned
        .grad) is copied in and re-pointed; a missing gradient counts as zero, like a parameter torch's SGD would skip
        ... except that momentum still decays -- so such parameters are not supported silently: they raise."""
        a = self._arena
        for g in self.param_groups:
            for p in g["params"]:
                o, n = a["offs"][id(p)], p.numel()
                view = a["g"][o:o + n].view(p.shape)
                if p.grad is None:
                    raise RuntimeError("ArenaSGD: a parameter has no gradient (torch's SGD would skip it); freeze it with "
                                       "requires_grad=False before building the optimizer instead")
                if p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                    p.grad = view

    # ---- torch.optim.Optimizer interface ------------------------------------------------------------------
    def zero_grad(self, set_to_none: bool = True) -> None:
        """Clears the gradient arena in place with one memset (gradients stay views of the arena, never None); a no-op
        when the step kernel resets them (`zero_grad_in_step=True`)."""
        if self._arena is None:
            return super().zero_grad(set_to_none=False)
        if not self.zero_grad_in_step:
            self._arena["g"].zero_()

    @torch.no_grad()
    def step(self, closure: Optional[Callable] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._arena is None or self._arena["sig"] != self._signature():
            self._build()
        a = self._arena
        self._adopt_stray_gradients()
        # per-group hyper-parameters -> device (fp32, like torch casts the python scalars): async copy from a small ring
        # of pinned rows, skipped when nothing changed since the last step
        cur = tuple((float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]), float(g["dampening"]))
                    for g in self.param_groups)
        if cur != a["last"]:
            k = a["tick"] % len(a["staging"])
            if a["events"][k] is not None:
                a["events"][k].synchronize()  # the copy that last used this row has run (8 steps ago: never blocks)
            a["staging"][k].copy_(torch.tensor(cur, dtype=torch.float64).to(torch.float32))
            a["hyper"].copy_(a["staging"][k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(a["dev"]))
            a["events"][k] = ev
            a["tick"] += 1
            a["last"] = cur
        L.check(L.lib().kodb_sgd_step(a["p"].data_ptr(), a["g"].data_ptr(), a["m"].data_ptr(), a["seg"],
                                      len(self.param_groups), a["hyper"].data_ptr(), int(a["nesterov"]),
                                      int(self.zero_grad_in_step), H.stream_ptr(a["dev"])), "kodb_sgd_step")
        return loss

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        if self._arena is not None:  # loaded momentum buffers are fresh tensors: move them back into the arena
            self._arena = None
            self._build()


class SmartOptimizer(object):
    """kod/nn/optim/smart.py:11-60.  `optimizer` is a factory (e.g. `partial(ArenaSGD, lr=0.01, momentum=0.937,
    nesterov=True)`, or torch's SGD); groups: biases (no decay), other parameters (weight_decay), norm weights (no decay)."""

    def __init__(self, optimizer: Callable[..., torch.optim.Optimizer], weight_decay: float):
        self.partial_opt = optimizer
        self.weight_decay = weight_decay

    def __call__(self, net: nn.Module) -> torch.optim.Optimizer:
        norm_types = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
        bias, decay, no_decay = [], [], []
        for module in net.modules():
            for name, p in module.named_parameters(recurse=False):
                if name == "bias":
                    bias.append(p)
                elif name == "weight" and isinstance(module, norm_types):
                    no_decay.append(p)
                else:
                    decay.append(p)
        opt = self.partial_opt(params=bias)
        opt.param_groups[0]["name"] = "bias_params"
        opt.add_param_group(dict(params=decay, weight_decay=self.weight_decay, name="decay_params"))
        opt.add_param_group(dict(params=no_decay, weight_decay=0.0, name="norm_params"))
        return opt


class OptimizerWarmupUpdater(object):
    """kod/lightning/experiments/yv5_baseline/warmup.py:11-58."""

    def __init__(self, warmup_epochs: int, warmup_bias_lr: float, warmup_momentum: float, momentum: Optional[float] = None):
        self.warmup_bias_lr = warmup_bias_lr
        self.warmup_momentum = warmup_momentum
        self.warmup_epochs = warmup_epochs
        self.momentum = momentum

    def __call__(self, current_step: int, current_epoch: int, max_warmup_steps: int, sch_fn: Callable,
                 optimizer: torch.optim.Optimizer) -> None:
        xi = [0, max_warmup_steps]
        for pg in optimizer.param_groups:
            start = self.warmup_bias_lr if pg["name"] == "bias_params" else 0.0
            pg["lr"] = np.interp(current_step, xi, [start, pg["initial_lr"] * sch_fn(current_epoch)])
            if "momentum" in pg:
                assert self.momentum is not None
                pg["momentum"] = np.interp(current_step, xi, [self.warmup_momentum, self.momentum])
