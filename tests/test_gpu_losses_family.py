"""Dice-family losses beyond DiceLoss (csrc/segloss.cu) against golden values and gradients produced by the reference classes
themselves (tests/golden/make_golden.py::losses_case: torch_em/loss/dice.py:136-256, combined_loss.py, distance_based.py).
fp32 tolerance: rtol 1e-5 on the loss, rtol 1e-4 / atol 1e-7 on the gradient (reduction order only)."""
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "losses.npz"))


def _run(loss_fn, inp, tgt):
    inp = torch.from_numpy(inp).to(DEV).requires_grad_(True)
    l = loss_fn(inp, torch.from_numpy(tgt).to(DEV))
    (l.sum() if l.dim() > 0 else l).backward()
    return l.detach().cpu().numpy(), inp.grad.cpu().numpy()


CASES = {
    "dice_logits_sum": (lambda: tb.DiceLossWithLogits(reduce_channel="sum"), "x", "t"),
    "dice_logits_mean": (lambda: tb.DiceLossWithLogits(reduce_channel="mean"), "x", "t"),
    "dice_logits_max": (lambda: tb.DiceLossWithLogits(reduce_channel="max"), "x", "t"),
    "dice_logits_min": (lambda: tb.DiceLossWithLogits(reduce_channel="min"), "x", "t"),
    "dice_logits_None": (lambda: tb.DiceLossWithLogits(reduce_channel=None), "x", "t"),
    "dice_logits_pooled": (lambda: tb.DiceLossWithLogits(channelwise=False), "x", "t"),
    "bce_dice": (lambda: tb.BCEDiceLoss(alpha=0.7, beta=1.3), "p", "t"),
    "bce_dice_pooled": (lambda: tb.BCEDiceLoss(alpha=1.0, beta=0.5, channelwise=False), "p", "t"),
    "bce_dice_logits": (lambda: tb.BCEDiceLossWithLogits(alpha=0.7, beta=1.3), "x", "t"),
    "combined": (lambda: tb.CombinedLoss(tb.DiceLoss(), tb.BCEDiceLoss(), loss_weights=[0.25, 0.75]), "p", "t"),
    "distance_True": (lambda: tb.DistanceLoss(mask_distances_in_bg=True), "p", "td"),
    "distance_False": (lambda: tb.DistanceLoss(mask_distances_in_bg=False), "p", "td"),
    "dice_distance_True": (lambda: tb.DiceBasedDistanceLoss(mask_distances_in_bg=True), "p", "td"),
    "dice_distance_False": (lambda: tb.DiceBasedDistanceLoss(mask_distances_in_bg=False), "p", "td"),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_loss_matches_reference_golden(z, name):
    make, ki, kt = CASES[name]
    loss, grad = _run(make(), z[ki], z[kt])
    np.testing.assert_allclose(loss, z["loss_" + name], rtol=1e-5)
    g = z["grad_" + name]
    np.testing.assert_allclose(grad, g, rtol=1e-4, atol=1e-7 + 1e-6 * np.abs(g).max())


def test_distance_loss_with_custom_members_keeps_reference_semantics(z):
    """Member losses the fused kernel does not know (L1 on the distances) run literally as distance_based.py:34-57."""
    p, td = torch.from_numpy(z["p"]).to(DEV).requires_grad_(True), torch.from_numpy(z["td"]).to(DEV)
    l = tb.DistanceLoss(True, foreground_loss=tb.DiceLoss(), distance_loss=torch.nn.L1Loss())(p, td)
    fg = td[:, :1]
    ref = tb.DiceLoss()(p[:, :1], fg) + sum(torch.nn.functional.l1_loss(p[:, c:c + 1] * fg, td[:, c:c + 1] * fg) for c in (1, 2))
    assert abs(l.item() - ref.item()) < 1e-6
    l.backward()
    assert p.grad is not None and bool(torch.isfinite(p.grad).all())


def test_bce_clamps_like_aten():
    """p exactly 0 / 1: F.binary_cross_entropy clamps log at -100 and the gradient denominator at 1e-12."""
    p = torch.tensor([0.0, 1.0, 0.0, 1.0, 0.5], device=DEV).reshape(1, 1, 5).requires_grad_(True)
    t = torch.tensor([0.0, 1.0, 1.0, 0.0, 1.0], device=DEV).reshape(1, 1, 5)
    l = tb.BCEDiceLoss(alpha=0.0, beta=1.0)(p, t)
    l.backward()
    pr = p.detach().clone().requires_grad_(True)
    lr = torch.nn.functional.binary_cross_entropy(pr, t)
    lr.backward()
    np.testing.assert_allclose(l.item(), lr.item(), rtol=1e-6)
    np.testing.assert_allclose(p.grad.cpu().numpy(), pr.grad.cpu().numpy(), rtol=1e-5)


def test_losses_accept_bf16_prediction_and_validate_shapes(z):
    x, t = torch.from_numpy(z["x"]).to(DEV), torch.from_numpy(z["t"]).to(DEV)
    xb = x.bfloat16().requires_grad_(True)
    l = tb.BCEDiceLossWithLogits()(xb, t)
    l.backward()
    assert xb.grad.dtype == torch.bfloat16
    ref = tb.BCEDiceLossWithLogits()(xb.detach().float(), t)
    assert abs(l.item() - ref.item()) < 1e-5 * abs(ref.item())
    with pytest.raises(ValueError, match="same shape"):
        tb.DiceLossWithLogits()(x, t[:, :2])
    with pytest.raises(AssertionError):
        tb.DistanceLoss()(x[:, :2], t[:, :2])
    with pytest.raises(RuntimeError, match="CUDA"):
        tb.BCEDiceLoss()(torch.rand(1, 1, 4), torch.rand(1, 1, 4))
