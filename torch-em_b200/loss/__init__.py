from .affinity import AffinityLoss
from .combined_loss import CombinedLoss
from .dice import BCEDiceLoss, BCEDiceLossWithLogits, DiceLoss, DiceLossWithLogits, dice_score
from .distance_based import DiceBasedDistanceLoss, DistanceLoss
from .wrapper import ApplyAndRemoveMask, ApplyMask, LossWrapper, MaskIgnoreLabel

__all__ = ["AffinityLoss", "DiceLoss", "DiceLossWithLogits", "BCEDiceLoss", "BCEDiceLossWithLogits", "CombinedLoss", "DistanceLoss",
           "DiceBasedDistanceLoss", "dice_score", "ApplyAndRemoveMask", "ApplyMask", "LossWrapper", "MaskIgnoreLabel"]
