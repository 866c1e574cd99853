"""Oracle: Dice score / loss and the mask transforms (TEST INFRASTRUCTURE ONLY).

Restates, in plain torch fp32 (autograd gives the backward) and in closed form (``dice_grad``):
  * flatten_samples + dice_score + DiceLoss          loss/dice.py:7-31, 34-93, 96-133
  * ApplyAndRemoveMask("multiply") / _multiply        loss/wrapper.py:84-87, 129-152
  * MaskIgnoreLabel("multiply")                       loss/wrapper.py:155-183
"""
import torch


def dice_score(input_, target, invert=False, channelwise=True, reduce_channel="sum", eps=1e-7):
    if input_.shape != target.shape:
        raise ValueError(f"Expect input and target of same shape, got: {input_.shape}, {target.shape}.")
    if channelwise:
        c = input_.shape[1]
        p = input_.transpose(0, 1).reshape(c, -1)
        t = target.transpose(0, 1).reshape(c, -1)
        num = (p * t).sum(-1)
        den = (p * p).sum(-1) + (t * t).sum(-1)
        score = 2 * (num / den.clamp(min=eps))
        if invert:
            score = 1.0 - score
        if reduce_channel is None:
            return score
        if reduce_channel in ("sum", "mean", "max", "min"):
            return getattr(score, reduce_channel)()
        raise ValueError(f"Unsupported channel reduction {reduce_channel}")
    num = (input_ * target).sum()
    den = (input_ * input_).sum() + (target * target).sum()
    score = 2.0 * (num / den.clamp(min=eps))
    return 1.0 - score if invert else score


def dice_loss(input_, target, channelwise=True, eps=1e-7, reduce_channel="sum"):
    return dice_score(input_, target, invert=True, channelwise=channelwise, eps=eps, reduce_channel=reduce_channel)


def apply_and_remove_mask_multiply(prediction, target):
    """wrapper.py:145-152 with masking_method='multiply'."""
    assert target.dim() == prediction.dim(), f"{target.dim()}, {prediction.dim()}"
    assert target.size(1) == 2 * prediction.size(1), f"{target.size(1)}, {prediction.size(1)}"
    assert target.shape[2:] == prediction.shape[2:], f"{str(target.shape)}, {str(prediction.shape)}"
    c = target.size(1) // 2
    mask = target[:, c:]
    target = target[:, :c]
    return prediction * mask, target * mask


def masked_dice_loss(prediction, target, **kw):
    """LossWrapper(DiceLoss(), ApplyAndRemoveMask('multiply')) -- the reference's affinity loss idiom."""
    p, t = apply_and_remove_mask_multiply(prediction, target)
    return dice_loss(p, t, **kw)


def ignore_label_dice_loss(prediction, target, ignore_label=-1, **kw):
    """LossWrapper(DiceLoss(), MaskIgnoreLabel(ignore_label, 'multiply'))."""
    mask = (target != ignore_label).to(prediction.dtype)
    return dice_loss(prediction * mask, target * mask, **kw)


def dice_grad(prediction, target, mask=None, eps=1e-7):
    """Closed-form d(loss)/d(prediction) for the channelwise / reduce='sum' loss."""
    p = prediction.double()
    t = target.double()
    m = torch.ones_like(p) if mask is None else mask.double()
    pm, tm = p * m, t * m
    dims = [0] + list(range(2, p.dim()))
    num = (pm * tm).sum(dims, keepdim=True)
    den = (pm * pm).sum(dims, keepdim=True) + (tm * tm).sum(dims, keepdim=True)
    live = den > eps
    a = torch.where(live, -2.0 / den.clamp(min=eps), torch.full_like(den, -2.0 / eps))
    b = torch.where(live, 4.0 * num / den.clamp(min=eps) ** 2, torch.zeros_like(den))
    return ((a * tm + b * pm) * m).to(prediction.dtype)
