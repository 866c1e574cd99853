"""``DistanceLoss`` / ``DiceBasedDistanceLoss`` with torch-em's signatures (torch_em/loss/distance_based.py:7-69).

Three channels: foreground + two distance channels.  With the default member losses (Dice on the foreground, mean-squared
error -- or Dice -- on the distances, optionally masked by the foreground TARGET) the whole loss is one fused reduction pass
and one backward pass (csrc/segloss.cu); other member losses keep the reference semantics literally.
"""
import torch
import torch.nn as nn

from .dice import DiceLoss
from .segloss import SegLossFn


class DistanceLoss(nn.Module):
    def __init__(self, mask_distances_in_bg: bool = True, foreground_loss: nn.Module = None,
                 distance_loss: nn.Module = None) -> None:
        super().__init__()
        self.foreground_loss = DiceLoss() if foreground_loss is None else foreground_loss
        self.distance_loss = nn.MSELoss(reduction="mean") if distance_loss is None else distance_loss
        self.mask_distances_in_bg = mask_distances_in_bg
        self.init_kwargs = {"mask_distances_in_bg": mask_distances_in_bg}

    def _fused_config(self):
        fg, ds = self.foreground_loss, self.distance_loss
        if not (type(fg) is DiceLoss and fg.channelwise and fg.reduce_channel in ("sum", "mean", "max", "min")):
            return None
        m = 1.0 if self.mask_distances_in_bg else 0.0
        if type(ds) is nn.MSELoss and ds.reduction == "mean":
            return [[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, m], [0.0, 0.0, 1.0, m]], fg.eps
        if type(ds) is DiceLoss and ds.channelwise and ds.eps == fg.eps and ds.reduce_channel in ("sum", "mean", "max", "min"):
            # each member Dice sees ONE channel, so every channel reduction is the identity: the sum of three Dice errors
            return [[1.0, 0.0, 0.0, 0.0], [1.0, 0.0, 0.0, m], [1.0, 0.0, 0.0, m]], fg.eps
        return None

    def forward(self, input_: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        assert input_.shape == target.shape, input_.shape
        assert input_.shape[1] == 3, input_.shape
        cfg = self._fused_config()
        if cfg is not None:
            chan, eps = cfg
            return SegLossFn.apply(input_, target, chan, False, 0 if self.mask_distances_in_bg else None, True, eps, "sum")
        # reference semantics, literally (distance_based.py:34-57); the single-channel slices keep their channel axis
        fg_input, fg_target = input_[:, 0:1], target[:, 0:1]
        overall = self.foreground_loss(fg_input, fg_target)
        for c in (1, 2):
            d_input, d_target = input_[:, c:c + 1], target[:, c:c + 1]
            if self.mask_distances_in_bg:
                overall = overall + self.distance_loss(d_input * fg_target, d_target * fg_target)
            else:
                overall = overall + self.distance_loss(d_input, d_target)
        return overall


class DiceBasedDistanceLoss(DistanceLoss):
    def __init__(self, mask_distances_in_bg: bool) -> None:
        super().__init__(mask_distances_in_bg, foreground_loss=DiceLoss(), distance_loss=DiceLoss())
