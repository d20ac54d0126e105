"""GPU: the optimizer step + warm-up -- SmartOptimizer + ArenaSGD (kodb_sgd_step) +
OptimizerWarmupUpdater -- against the golden produced by the real reference classes (torch.optim.SGD nesterov)."""
from functools import partial

import numpy as np
import pytest
import torch
import torch.nn as nn

import _golden as G
from object_detection_cib_b200.optim import ArenaSGD, OptimizerWarmupUpdater, SmartOptimizer
from oracle import sgd as osgd

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net_from_golden(g):
    net = nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8), nn.SiLU(), nn.Conv2d(8, 5, 1, bias=True),
                        nn.GroupNorm(1, 5), nn.Flatten(), nn.Linear(80, 7)).to(DEV)
    with torch.no_grad():
        for n, p in net.named_parameters():
            p.copy_(torch.from_numpy(g[f"init/{n}"]))
    return net


@pytest.mark.parametrize("fused_zero", [False, True])
def test_sgd_warmup_golden(fused_zero):
    g = G.load("sgd_warmup")
    net = _net_from_golden(g)
    names = [str(n) for n in g["names"]]
    params = dict(net.named_parameters())
    opt = SmartOptimizer(partial(ArenaSGD, lr=0.01, momentum=0.937, nesterov=True, zero_grad_in_step=fused_zero),
                         weight_decay=0.0005)(net)
    assert [pg["name"] for pg in opt.param_groups] == [str(n) for n in g["group_names"]]
    group_of = {id(q): gi for gi, pg in enumerate(opt.param_groups) for q in pg["params"]}
    assert [group_of[id(params[n])] for n in names] == g["group_index"].tolist()  # same grouping as smart.py:24-35
    for pg in opt.param_groups:
        pg.setdefault("initial_lr", pg["lr"])
    upd = OptimizerWarmupUpdater(warmup_epochs=3, warmup_bias_lr=0.1, warmup_momentum=0.8, momentum=0.937)
    sch = partial(osgd.sch_linear, max_epochs=30, lrf=0.01)
    steps, nw = g["hyper"].shape[0], 5
    for step in range(steps):
        opt.zero_grad()
        for n in names:
            gr = torch.from_numpy(g[f"grad{step}/{n}"]).to(DEV)
            if step == 0 or step == 3:
                params[n].grad = gr  # a stray gradient tensor: adopted into the arena by step()
            else:
                assert float(params[n].grad.abs().max()) == 0.0  # cleared by zero_grad() / by the step kernel
                params[n].grad.add_(gr)  # what autograd's accumulation does
        if step <= nw:
            upd(current_step=step, current_epoch=step // 3, max_warmup_steps=nw, sch_fn=sch, optimizer=opt)
        got = [[float(pg["lr"]), float(pg["momentum"]), float(pg["weight_decay"])] for pg in opt.param_groups]
        assert np.array_equal(np.array(got), g["hyper"][step])
        opt.step()
    flat = opt._arena["p"]
    for n in names:
        p = params[n]
        assert flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + flat.numel() * 4  # parameters are arena views
        assert np.array_equal(p.detach().cpu().numpy().view(np.uint32), g[f"final/{n}"].view(np.uint32)), n
        m = opt.state[p]["momentum_buffer"]
        assert np.array_equal(m.cpu().numpy().view(np.uint32), g[f"mom/{n}"].view(np.uint32)), n
    # the network still runs on the re-pointed parameters, and a checkpoint round trip keeps the momentum
    net(torch.zeros(2, 3, 6, 6, device=DEV)).sum().backward()
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    m0 = opt.state[params[names[0]]]["momentum_buffer"]
    assert np.array_equal(m0.cpu().numpy().view(np.uint32), g[f"mom/{names[0]}"].view(np.uint32))


def test_sgd_matches_torch_sgd_on_a_larger_net():
    """~1.2 M parameters, 4 steps, against torch.optim.SGD on the same device (bit-exact: same fused multiply-adds)."""
    torch.manual_seed(0)
    mk = lambda: nn.Sequential(nn.Conv2d(16, 128, 3), nn.BatchNorm2d(128), nn.SiLU(), nn.Conv2d(128, 256, 3), nn.BatchNorm2d(256),
                               nn.SiLU(), nn.Conv2d(256, 333, 3, bias=True), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(333, 10)).to(DEV)
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    oa = SmartOptimizer(partial(ArenaSGD, lr=0.02, momentum=0.9, nesterov=True), weight_decay=5e-4)(a)
    ob = SmartOptimizer(partial(torch.optim.SGD, lr=0.02, momentum=0.9, nesterov=True), weight_decay=5e-4)(b)
    x = torch.randn(4, 16, 12, 12, device=DEV)
    for step in range(4):
        ob.zero_grad()
        b(x).square().mean().backward()
        oa.zero_grad()
        if step == 0:  # (first backward through `a` itself: autograd creates the gradients, the step adopts them)
            a(x).square().mean().backward()
        # both optimizers must see the SAME gradients: cuDNN's backward kernels are not bit-reproducible between calls
        for pa, pb in zip(a.parameters(), b.parameters()):
            pa.grad.copy_(pb.grad)
        oa.step()
        ob.step()
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa, pb), n


def test_sgd_refuses_cpu_parameters():
    net = nn.Linear(4, 4)
    opt = ArenaSGD(net.parameters(), lr=0.1, momentum=0.9)
    net(torch.zeros(1, 4)).sum().backward()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
