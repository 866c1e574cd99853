"""``LossWrapper`` and the mask transforms with torch-em's signatures (torch_em/loss/wrapper.py:7-183).

``LossWrapper(DiceLoss(), ApplyAndRemoveMask("multiply"))`` -- the reference's affinity-loss idiom
(torch_em/cli.py:263-267) -- and ``LossWrapper(DiceLoss(), MaskIgnoreLabel(..., "multiply"))`` are recognised and
run as ONE fused masked-Dice kernel pair (no ``prediction * mask`` / ``target * mask`` tensors).  Other combinations
keep the reference semantics literally: the transform is applied, then the loss.
"""
from typing import Callable, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .dice import DiceLoss, masked_dice

__all__ = ["LossWrapper", "ApplyMask", "ApplyAndRemoveMask", "MaskIgnoreLabel"]


def _crop(prediction, target, mask, channel_dim):
    if mask.shape[channel_dim] != 1:
        raise ValueError(
            "_crop only supports a mask with a singleton channel axis. Please consider using masking_method=multiply."
        )
    mask = mask.type(torch.bool).squeeze(channel_dim)
    return prediction.moveaxis(channel_dim, -1)[mask], target.moveaxis(channel_dim, -1)[mask]


def _multiply(prediction, target, mask, channel_dim):
    return prediction * mask, target * mask


class ApplyMask:
    """Mask prediction and target by 'crop' or 'multiply' (wrapper.py:90-126)."""
    MASKING_FUNCS = {"crop": _crop, "multiply": _multiply}

    def __init__(self, masking_method: str = "crop", channel_dim: int = 1):
        if masking_method not in self.MASKING_FUNCS.keys():
            raise ValueError(f"{masking_method} is not available, please use one of {list(self.MASKING_FUNCS.keys())}.")
        self.masking_method = masking_method
        self.masking_func = self.MASKING_FUNCS[masking_method]
        self.channel_dim = channel_dim
        self.init_kwargs = {"masking_method": masking_method, "channel_dim": channel_dim}

    def __call__(self, prediction: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        mask.requires_grad = False
        return self.masking_func(prediction, target, mask, self.channel_dim)


class ApplyAndRemoveMask(ApplyMask):
    """Take the mask from the second half of the target channels (wrapper.py:129-152)."""

    def __call__(self, prediction: torch.Tensor, target: torch.Tensor):
        assert target.dim() == prediction.dim(), f"{target.dim()}, {prediction.dim()}"
        assert target.size(1) == 2 * prediction.size(1), f"{target.size(1)}, {prediction.size(1)}"
        assert target.shape[2:] == prediction.shape[2:], f"{str(target.shape)}, {str(prediction.shape)}"
        seperating_channel = target.size(1) // 2
        mask = target[:, seperating_channel:]
        target = target[:, :seperating_channel]
        return super().__call__(prediction, target, mask)


class MaskIgnoreLabel(ApplyMask):
    """Mask where target == ignore_label (wrapper.py:155-183)."""

    def __init__(self, ignore_label: int = -1, masking_method: str = "crop", channel_dim: int = 1):
        super().__init__(masking_method, channel_dim)
        self.ignore_label = ignore_label
        self.init_kwargs["ignore_label"] = ignore_label

    def __call__(self, prediction: torch.Tensor, target: torch.Tensor):
        mask = (target != self.ignore_label)
        return super().__call__(prediction, target, mask)


class LossWrapper(nn.Module):
    """Apply a transform to prediction / target, then the loss (wrapper.py:7-62)."""

    def __init__(self, loss: nn.Module, transform: Callable):
        super().__init__()
        self.loss = loss
        if not callable(transform):
            raise ValueError("transform has to be callable.")
        self.transform = transform
        self.init_kwargs = {"loss": loss, "transform": transform}

    def _fused(self, prediction, target, kwargs):
        if kwargs or not isinstance(self.loss, DiceLoss) or not torch.is_tensor(prediction) or prediction.dim() < 3:
            return None
        tr, ls = self.transform, self.loss
        if type(tr) is ApplyAndRemoveMask and tr.masking_method == "multiply" and tr.channel_dim == 1:
            return masked_dice(prediction, target, ls.channelwise, ls.eps, ls.reduce_channel)
        if type(tr) is MaskIgnoreLabel and tr.masking_method == "multiply" and tr.channel_dim == 1:
            return ls.forward_masked(prediction, target, target != tr.ignore_label)
        return None

    def apply_transform(self, prediction, target, **kwargs):
        """@private"""
        if isinstance(prediction, (list, tuple)):
            assert isinstance(target, (list, tuple))
            out_p, out_t = [], []
            for pred, targ in zip(prediction, target):
                p, t = self.transform(pred, targ, **kwargs)
                out_p.append(p)
                out_t.append(t)
            return out_p, out_t
        return self.transform(prediction, target, **kwargs)

    def forward(self, prediction: Union[Sequence[torch.Tensor], torch.Tensor],
                target: Union[Sequence[torch.Tensor], torch.Tensor], **kwargs) -> torch.Tensor:
        fused = self._fused(prediction, target, kwargs)
        if fused is not None:
            return fused
        prediction, target = self.apply_transform(prediction, target, **kwargs)
        return self.loss(prediction, target)
