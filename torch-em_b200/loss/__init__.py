from .affinity import AffinityLoss
from .dice import DiceLoss, dice_score
from .wrapper import ApplyAndRemoveMask, ApplyMask, LossWrapper, MaskIgnoreLabel

__all__ = ["AffinityLoss", "DiceLoss", "dice_score", "ApplyAndRemoveMask", "ApplyMask", "LossWrapper", "MaskIgnoreLabel"]
