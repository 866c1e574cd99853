"""GPU parity of the whole network (forward, loss, every parameter gradient) against the golden vectors generated
from the reference itself, and of the bf16 path against the fp32 oracle.

Tolerances: fp32 -- the reference's own batched-vs-unbatched bound rtol 1e-4 / atol 1e-4 on outputs
(test/util/test_prediction.py:358-382), gradients rtol 2e-3 with an absolute floor of 1e-4 x the largest entry;
bf16 -- relative L2 error <= 2e-2 on the prediction, cosine similarity >= 0.99 per parameter gradient.
"""
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from oracle import dice as odice
from oracle import unet as ounet
from tests.test_engine_cpu import CASES, build

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", sorted(CASES))
def test_fp32_matches_reference_golden(golden_dir, name):
    z, net = build(golden_dir, name)
    net.to(DEV)
    x, t = torch.from_numpy(z["x"]).to(DEV), torch.from_numpy(z["t"]).to(DEV)
    y = net(x)
    assert y.dtype == torch.float32 and tuple(y.shape) == z["y"].shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), z["y"], rtol=1e-4, atol=1e-4)
    y.retain_grad()                                      # default_trainer.py:798-800
    loss = tb.DiceLoss()(y, t)
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-4)
    loss.backward()
    assert y.grad is not None
    for k, p in net.named_parameters():
        g = z["g:" + k]
        np.testing.assert_allclose(p.grad.cpu().numpy(), g, rtol=2e-3, atol=2e-6 + 1e-4 * np.abs(g).max(), err_msg=k)


@pytest.mark.parametrize("norm", ["InstanceNorm", "GroupNorm"])
def test_bf16_autocast_no_worse_than_reference_autocast(norm):
    """bf16 has no exact answer: the yardstick is the fp32 oracle, and the bar is the error the reference's OWN bf16
    autocast run (same functional graph through cuDNN/ATen on this GPU) makes against it.  Ours must be within
    1.25x of that error on the prediction; per parameter gradient the absolute error must be <= 2x the reference
    autocast's (+ a floor of 5e-3 of the largest gradient norm: we keep the network input in bf16, the
    reference's first norm sees it in fp32) and the median error ratio <= 1.1."""
    torch.manual_seed(0)
    kw = dict(in_channels=1, out_channels=2, depth=3, initial_features=16, final_activation="Sigmoid", norm=norm)
    net = tb.UNet3d(**kw).to(DEV)
    x = torch.randn(2, 1, 32, 32, 32)
    t = (torch.rand(2, 2, 32, 32, 32) > 0.5).float()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    y_ref = ounet.unet3d_forward(x, sd, [2, 2, 2], norm=norm, final_activation="Sigmoid")
    l_ref = odice.dice_loss(y_ref, t)
    l_ref.backward()
    sdg = {k: v.detach().to(DEV).clone().requires_grad_(True) for k, v in net.state_dict().items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y_ac = ounet.unet3d_forward(x.to(DEV), sdg, [2, 2, 2], norm=norm, final_activation="Sigmoid")
        odice.dice_loss(y_ac, t.to(DEV)).backward()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = net(x.to(DEV))
        loss = tb.DiceLoss()(y, t.to(DEV))
    loss.backward()

    def rel(a, b):
        return float((a.detach().float().cpu() - b.detach()).norm() / (b.detach().norm() + 1e-30))

    assert rel(y, y_ref) < 2e-2                                   # relative L2 <= 1e-2..2e-2
    assert rel(y, y_ref) <= 1.25 * rel(y_ac, y_ref) + 1e-3
    assert abs(loss.item() - l_ref.item()) < 1e-2 * abs(l_ref.item())
    # Per parameter: absolute error (L2) against the fp32 gradient, in units of the largest parameter-gradient norm.
    # (Relative error is meaningless for gradients that are differences of large terms -- e.g. a conv bias or a norm
    # scale sitting in front of conv -> norm is ~0 in exact arithmetic.)
    gmax = max(float(v.grad.norm()) for v in sd.values())

    def err(a, b):
        return float((a.detach().float().cpu() - b.detach()).norm()) / gmax

    ratios = {}
    for k, p in net.named_parameters():
        e_ours, e_ref = err(p.grad, sd[k].grad), err(sdg[k].grad, sd[k].grad)
        ratios[k] = (e_ours, e_ref)
        assert e_ours <= 2.0 * e_ref + 5e-3, (k, e_ours, e_ref)
    r = sorted(a / (b + 1e-4) for a, b in ratios.values())
    assert r[len(r) // 2] <= 1.1, ("median error ratio vs reference autocast", r[len(r) // 2])


def test_fp16_autocast_runs_on_the_h16_path():
    """torch.autocast(float16) + GradScaler -- the reference trainer's DEFAULT mixed precision (default_trainer.py:132-142) -- is
    served by the h16 path: fp32 activations, fp16 tensor-core operands with range scaling (so the scaler's 65536x loss scale is
    harmless), fp32 outputs and gradients.  At least fp16-autocast accuracy: checked against the fp32 oracle at TF32-class
    tolerance, whatever allow_tf32 says (it is False in this test suite)."""
    from oracle import dice as odice
    from oracle import unet as ounet
    from torch_em_b200.backend import default_backend
    torch.manual_seed(5)
    kw = dict(in_channels=1, out_channels=2, depth=2, initial_features=32, final_activation="Sigmoid")
    net = tb.UNet3d(**kw).to(DEV)
    x = torch.randn(1, 1, 16, 32, 32)
    t = (torch.nn.functional.avg_pool3d(torch.randn(1, 2, 16, 32, 32), 5, 1, 2) > 0).float()
    B = default_backend()
    B.calls.clear()
    scaler = torch.amp.GradScaler("cuda")
    with torch.autocast("cuda", dtype=torch.float16):
        y = net(x.to(DEV))
        loss = tb.DiceLoss()(y, t.to(DEV))
    scaler.scale(loss).backward()
    torch.cuda.synchronize()
    assert y.dtype == torch.float32
    assert B.calls.get("h16:fwd", 0) > 0 and B.calls.get("h16:dgrad", 0) > 0 and B.calls.get("h16:wgrad", 0) > 0, dict(B.calls)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    y_ref = ounet.unet3d_forward(x, sd, [2, 2], final_activation="Sigmoid")
    l_ref = odice.dice_loss(y_ref, t)
    l_ref.backward()
    assert float((y.detach().cpu() - y_ref.detach()).norm() / y_ref.detach().norm()) < 2e-3
    assert abs(loss.item() - l_ref.item()) < 2e-3 * abs(l_ref.item())
    s_ = scaler.get_scale()
    g = torch.cat([p.grad.flatten().cpu() / s_ for p in net.parameters()])
    gref = torch.cat([sd[k].grad.flatten() for k, _ in net.named_parameters()])
    assert bool(torch.isfinite(g).all())
    assert float(torch.dot(g, gref) / (g.norm() * gref.norm())) > 0.99
    assert float((g - gref).norm() / gref.norm()) < 0.15
    assert not B.h16_enabled()                       # the forcing ends with the node (allow_tf32 is False here)


def test_train_steps_reduce_loss_and_match_oracle_trajectory():
    """Five AdamW steps (torch_em/segmentation.py:543 defaults) through our model+loss vs the fp32 oracle."""
    torch.manual_seed(1)
    kw = dict(in_channels=1, out_channels=2, depth=2, initial_features=8, final_activation="Sigmoid")
    net = tb.UNet3d(**kw).to(DEV)
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    x = torch.randn(2, 1, 16, 32, 32)
    t = (torch.rand(2, 2, 16, 32, 32) > 0.7).float()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    opt_ref = torch.optim.AdamW(list(sd.values()), lr=1e-3)
    ours, ref = [], []
    for _ in range(5):
        opt.zero_grad()
        l = tb.DiceLoss()(net(x.to(DEV)), t.to(DEV))
        l.backward()
        opt.step()
        ours.append(l.item())
        opt_ref.zero_grad()
        lr = odice.dice_loss(ounet.unet3d_forward(x, sd, [2, 2], final_activation="Sigmoid"), t)
        lr.backward()
        opt_ref.step()
        ref.append(lr.item())
    assert ours[-1] < ours[0]
    np.testing.assert_allclose(ours, ref, rtol=2e-3)


def test_weight_update_through_data_is_seen():
    """In-place parameter updates that do not bump the tensor version (``p.data.mul_()``: EMA, clamping, old-style
    optimizers) must reach the packed bf16 operand images: they are rebuilt at the start of every pass."""
    torch.manual_seed(0)
    net = tb.UNet3d(1, 2, depth=2, initial_features=32, final_activation="Sigmoid").to(DEV)
    x = torch.randn(1, 1, 16, 16, 16, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y0 = net(x).clone()
        v0 = [p._version for p in net.parameters()]
        for p in net.parameters():
            p.data.mul_(0.5)                                   # invisible to p._version
        assert [p._version for p in net.parameters()] == v0
        y1 = net(x).clone()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    y_ref = ounet.unet3d_forward(x.cpu(), sd, [2, 2], final_activation="Sigmoid")
    assert float((y1 - y0).abs().max()) > 1e-3                   # the output moved ...
    assert float((y1.cpu() - y_ref).norm() / y_ref.norm()) < 2e-2   # ... to where the NEW weights put it


def test_replica_on_moved_storage():
    """``model.to()`` / ``load_state_dict`` / ``copy.deepcopy`` (predict_with_halo, prediction.py:188-192): the pack set
    follows the parameters' storage and is not shared between copies."""
    import copy
    torch.manual_seed(0)
    net = tb.UNet3d(1, 2, depth=2, initial_features=32, final_activation="Sigmoid").to(DEV)
    x = torch.randn(1, 1, 16, 16, 16, device=DEV)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y0 = net(x).clone()
        twin = copy.deepcopy(net)
        assert "_pack_store" not in twin.__dict__
        for p in twin.parameters():
            p.data.zero_()
        yt = twin(x).clone()
        y1 = net(x).clone()
        net.cpu().to(DEV)                                       # new storage, same values
        y2 = net(x).clone()
    assert float((y1 - y0).abs().max()) < 1e-2 and float((y2 - y0).abs().max()) < 1e-2     # (fp32 atomics in the statistics)
    assert float((yt - 0.5).abs().max()) < 1e-6                   # all-zero weights: sigmoid(0)
